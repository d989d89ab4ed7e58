"""One evidence-path workload (2048 datasets x 1000 trials, 200 observations) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import basic_ddm_dc_evidence as mev
sim = pkg.DDMSimulator(0, seed=1)
Pe = mev.batch_draw_prior(2048)
for _ in range(3):
    b = sim.simulate_evidence(Pe, 1000, 200, 1, flags=2, device=True); del b
sim.synchronize()
print(sim.last_stats())
