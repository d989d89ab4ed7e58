"""BASELINE config 3 at scale for profiling: single_trial_alpha_not_scaled, 20 000 datasets x 1000 trials, dt = .01."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors
sim = pkg.DDMSimulator(0, seed=2023)
P = priors.draw_prior_batch("alpha", 20000, np.random.default_rng(2023))
sim._check(sim._lib.ddm_upload_params(sim._ctx, 1, P.ctypes.data_as(pkg._capi._dp), 20000, 7))
for _ in range(5):
    sim._check(sim._lib.ddm_run(sim._ctx, 1000, 0.01, 400, 2023, 0, 32, 2))
st = sim.last_stats()
print(f"kernel_ms {st['kernel_ms']:.3f} steps {st['total_steps']} steps/s {st['total_steps']/st['kernel_ms']*1e3:.3e} thr {st['refill_threshold']}")
