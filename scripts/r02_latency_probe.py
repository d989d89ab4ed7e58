"""Small launches (one trial per lane or fewer): the tile kernel against the one-thread-per-trial kernel, kernel time."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import _capi as capi, priors
sim = pkg.DDMSimulator(0, seed=2023)
def best(model, P, n, dt, ms, flags, reps=7):
    b = None
    for _ in range(reps):
        sim.run(model, P, n, dt, ms, flags=flags, dataset_offset=0)
        st = sim.last_stats()
        b = st if b is None or st["kernel_ms"] < b["kernel_ms"] else b
    return b
for model, prior, dt, ms in ((0, "basic", 0.01, 400), (0, "basic", 0.001, 4000), (1, "alpha", 0.01, 400)):
    for B, n in ((1, 300), (64, 500), (40, 500), (128, 500), (256, 500), (512, 500), (1024, 500), (1024, 1000), (4096, 1000)):
        P = priors.draw_prior_batch(prior, B, np.random.default_rng(3))
        sim.set_kernel_variant(0)
        t = best(model, P, n, dt, ms, capi.FLAG_OUT_F32)
        g = best(model, P, n, dt, ms, capi.FLAG_OUT_F32 | capi.FLAG_FORCE_GENERIC)
        sim.set_kernel_variant(2)
        l = best(model, P, n, dt, ms, capi.FLAG_OUT_F32)
        sim.set_kernel_variant(-1)
        same = t["total_steps"] == g["total_steps"] == l["total_steps"] and l["scheduler"] == 3 and t["scheduler"] == 2
        print(f"model {model} dt {dt} {B:5d} x {n:4d} = {B*n:8d} trials: tile {t['kernel_ms']*1e3:8.1f} us (grid {t['grid']})   one-thread-per-trial {g['kernel_ms']*1e3:8.1f} us   "
              f"latency kernel {l['kernel_ms']*1e3:8.1f} us   tile/latency {t['kernel_ms']/l['kernel_ms']:.2f}  same steps {same}", flush=True)
