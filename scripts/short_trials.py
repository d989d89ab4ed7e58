"""Throughput on the reference-default step (dt = .01, max_steps = 400: ~28 steps per trial), where the
finish/refill path weighs most, for the basic and the per-trial-boundary model."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors

sim = pkg.DDMSimulator(0, seed=2023)
for model, prior, D in ((0, "basic", 200_000), (1, "alpha", 200_000)):
    params = priors.draw_prior_batch(prior, D, np.random.default_rng(1))
    sim._check(sim._lib.ddm_upload_params(sim._ctx, model, params.ctypes.data_as(pkg._capi._dp), D, params.shape[1]))
    for thr in (0, 5, 8, 12, 16):
        sim.set_tuning(thr, 0, 0)
        best = None
        for _ in range(3):
            sim._check(sim._lib.ddm_run(sim._ctx, 1000, 0.01, 400, 2023, 0, 32, 2))
            st = sim.last_stats()
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        print(f"model {model} thr={thr:2d} ms={best['kernel_ms']:.3f} trials/s={D*1000/best['kernel_ms']*1e3:.3e} "
              f"steps/s={best['total_steps']/best['kernel_ms']*1e3:.3e} steps/trial={best['total_steps']/(D*1000):.1f}", flush=True)
