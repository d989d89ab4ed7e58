import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import basic_ddm_dc_evidence as mev
sim = pkg.DDMSimulator(0, seed=1)
for B in (256, 2048, 16384):
    Pe = mev.batch_draw_prior(B)
    for _ in range(3):
        b = sim.simulate_evidence(Pe, 1000, 200, 1, flags=2, device=True); del b
    sim.synchronize()
    st = sim.last_stats()
    print(B, "kernel_ms", st["kernel_ms"], "steps", st["total_steps"], "grid", st["grid"], flush=True)
