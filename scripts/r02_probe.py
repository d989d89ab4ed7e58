"""One configuration of the stepping kernel, for ncu: python scripts/r02_probe.py CASE VARIANT THR TILE [REPS]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors

case, variant, thr, tile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
CASES = {"sweep": (0, "sweep", 100_000, 1000, 1e-3, 4000), "c3": (1, "alpha", 20_000, 1000, 0.01, 400),
         "basic01": (0, "basic", 40_000, 1000, 0.01, 400), "sweep20k": (0, "sweep", 20_000, 1000, 1e-3, 4000),
         "c3small": (1, "alpha", 1024, 1000, 0.01, 400)}
model, prior, B, n, dt, ms = CASES[case]
sim = pkg.DDMSimulator(device=0, seed=2023)
sim.set_kernel_variant(variant)
sim.set_tuning(thr, 0, tile)
params = priors.draw_prior_batch(prior, B, np.random.default_rng(1))
for _ in range(reps):
    sim.run(model, params, n, dt, ms, seed=7, dataset_offset=0, flags=2)
    st = sim.last_stats()
print(case, "variant", variant, "thr", st["refill_threshold"], "tile", st["tile"], "kernel_ms", st["kernel_ms"],
      "steps/s %.4g" % (st["total_steps"] / st["kernel_ms"] * 1e3))
sim.close()
