"""C2 (64 x 500 training batch): ddm_training_batch against the three-call sequence."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import _capi as capi
sim = pkg.DDMSimulator(0, seed=2023)
def three():
    pd = sim.draw_prior("basic", 64); sim.run_uploaded(500, 0.01, 400, flags=capi.FLAG_OUT_F32); b = sim.last_output_dlpack(); del b
def one():
    pd, b = sim.training_batch("basic", 64, 500, 0.01, 400); del b
for name, f in (("three calls", three), ("ddm_training_batch", one), ("three calls", three), ("ddm_training_batch", one)):
    for _ in range(20): f()
    sim.synchronize()
    bl = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(200): f()
        bl.append((time.perf_counter() - t0) / 200 * 1e3)
    print(f"{name:20s} ms per batch: median {np.median(bl):.4f}  blocks {['%.4f' % x for x in bl]}  kernel {sim.last_stats()['kernel_ms']:.4f}", flush=True)
