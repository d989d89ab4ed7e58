"""Where does a C2 training batch (64 x 500, device prior -> simulate -> DLPack) spend its time?"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import _capi as capi
from bayesflow_nddms_b200 import priors

sim = pkg.DDMSimulator(device=0, seed=2023)


def loop(label, variant, reps=200):
    sim.set_kernel_variant(variant)
    for _ in range(10):
        sim.draw_prior("basic", 64)
        sim.run_uploaded(500, 0.01, 400, flags=capi.FLAG_OUT_F32)
    sim.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        sim.draw_prior("basic", 64)
        sim.run_uploaded(500, 0.01, 400, flags=capi.FLAG_OUT_F32)
        b = sim.last_output_dlpack()
        del b
    lat = (time.perf_counter() - t0) / reps
    st = sim.last_stats()
    print(f"{label:40s} variant {variant}: {lat * 1e3:.4f} ms/batch, kernel {st['kernel_ms']:.4f} ms, grid {st['grid']}, tile {st['tile']}, thr {st['refill_threshold']}", flush=True)


loop("fresh context", 0)
loop("fresh context", 1)
P = priors.draw_prior_batch("sweep", 100_000, np.random.default_rng(1))
sim.run(0, P, 1000, 1e-3, 4000, flags=2)
sim.synchronize()
loop("after a 1e8-trial resident run", 0)
loop("after a 1e8-trial resident run", 1)
out = sim.simulate(0, P[:20000], 1000, 1e-3, 4000, flags=2)
loop("after a streamed host run", 0)
sim.run(0, P[:2000], 1000, 1e-3, 4000, precision=64)
sim.synchronize()
loop("after an fp64 run", 0)
h = sim.host_stream_peak(0, 1 << 28)
loop("after host_stream_peak", 0)
sim.close()
