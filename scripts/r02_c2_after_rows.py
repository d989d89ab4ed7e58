"""Reproduce bench.py's state before its C2 leg (big streamed host-row runs into pinned arrays) and time the C2 loop's phases."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch

import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import _capi as capi
from bayesflow_nddms_b200 import basic_ddm_dc, priors

sim = pkg.DDMSimulator(device=0, seed=2023)
use_torch_stream = len(sys.argv) > 1 and sys.argv[1] == "torchstream"
if use_torch_stream:
    stream = torch.cuda.Stream(device=torch.device("cuda", 0))
    sim.set_stream(stream.cuda_stream)


def c2(label):
    for _ in range(10):
        sim.draw_prior("basic", 64)
        sim.run_uploaded(500, 0.01, 400, flags=capi.FLAG_OUT_F32)
    sim.synchronize()
    tp = tr = td = 0.0
    for _ in range(200):
        t0 = time.perf_counter()
        sim.draw_prior("basic", 64)
        t1 = time.perf_counter()
        sim.run_uploaded(500, 0.01, 400, flags=capi.FLAG_OUT_F32)
        t2 = time.perf_counter()
        b = sim.last_output_dlpack()
        del b
        t3 = time.perf_counter()
        tp += t1 - t0
        tr += t2 - t1
        td += t3 - t2
    print(f"{label:36s} prior {tp / 200 * 1e3:.4f}  run {tr / 200 * 1e3:.4f}  dlpack+del {td / 200 * 1e3:.4f} ms", flush=True)


c2("fresh")
D = 200_000
pe = priors.draw_prior_batch("sweep", D, np.random.default_rng(1))
h = basic_ddm_dc.batch_simulate_histogram(pe, 1000, sim, dt=1e-3, max_steps=4000, dataset_offset=0, n_bins=401, rt_max=4.01)
c2("after histogram e2e")
for label, dtype, fl in (("float64", np.float64, 0), ("float32", np.float32, 2)):
    out_host = sim.pinned_empty((D, 1000, 2), dtype, slot="rows_" + label)
    basic_ddm_dc.batch_simulate_trials(pe, 1000, sim, dt=1e-3, max_steps=4000, dataset_offset=0, out=out_host, flags=fl)
    del out_host
    c2("after host rows " + label)
sim.host_stream_peak(0, 1 << 30)
c2("after host_stream_peak")
sim.close()
