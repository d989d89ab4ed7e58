"""C2-sized launch (64 x 500, dt = .01) under the latency kernel (variant -1) or the tile kernel (variant 0), for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import _capi as capi, priors
sim = pkg.DDMSimulator(0, seed=2023)
sim.set_kernel_variant(int(sys.argv[1]) if len(sys.argv) > 1 else -1)
P = priors.draw_prior_batch("basic", 64, np.random.default_rng(3))
for _ in range(5):
    sim.run(0, P, 500, 0.01, 400, flags=capi.FLAG_OUT_F32, dataset_offset=0)
sim.synchronize()
print(sim.last_stats())
