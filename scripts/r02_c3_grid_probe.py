import sys
import numpy as np
sys.path.insert(0, ".")
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors
sim = pkg.DDMSimulator(device=0, seed=2023)
params = priors.draw_prior_batch("alpha", 1024, np.random.default_rng(1))
for bps in (0, 2, 3, 4):
    for tile in (32, 64, 128):
        for thr in (6, 8):
            sim.set_tuning(thr, bps, tile)
            best = 1e9
            for _ in range(8):
                sim.run(1, params, 1000, 0.01, 400, seed=7, dataset_offset=0, flags=2)
                st = sim.last_stats(); best = min(best, st["kernel_ms"])
            print("bps", bps, "tile", tile, "thr", thr, "grid", st["grid"], "kernel_ms %.4f" % best, flush=True)
