"""A/B on one B200: the round-1 persistent kernel (variant 1) against the tile kernel (variant 0), over refill
thresholds and tile sizes, device-resident float32 rows.  Prints one JSON line per configuration."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import priors

F32 = 2
sim = pkg.DDMSimulator(device=0, seed=2023)


def run(model, params, n, dt, ms, variant, thr, tile, reps=3):
    sim.set_kernel_variant(variant)
    sim.set_tuning(thr, 0, tile)
    best, st = 1e30, None
    for _ in range(reps + 1):
        sim.run(model, params, n, dt, ms, seed=7, dataset_offset=0, flags=F32)
        st = sim.last_stats()
        best = min(best, st["kernel_ms"])
    return best, st


def checksum(model, params, n, dt, ms, variant):
    sim.set_kernel_variant(variant)
    sim.set_tuning(0, 0, 0)
    out = sim.simulate(model, params[:2000], n, dt, ms, seed=7, dataset_offset=0, flags=F32)
    return float(out.astype(np.float64).sum()), sim.last_stats()["total_steps"]


cases = [
    ("sweep basic dt=.001", 0, priors.draw_prior_batch("sweep", 100_000, np.random.default_rng(1)), 1000, 1e-3, 4000,
     [(1, 5, 64), (0, 1, 128), (0, 2, 128), (0, 3, 128), (0, 4, 128), (0, 5, 128), (0, 6, 128), (0, 8, 128), (0, 3, 64)]),
    ("C3 alpha dt=.01", 1, priors.draw_prior_batch("alpha", 20_000, np.random.default_rng(2)), 1000, 0.01, 400,
     [(1, 16, 64), (0, 1, 128), (0, 2, 128), (0, 3, 128), (0, 4, 128), (0, 6, 128), (0, 8, 128), (0, 10, 128), (0, 12, 128), (0, 4, 64)]),
    ("basic dt=.01", 0, priors.draw_prior_batch("basic", 40_000, np.random.default_rng(3)), 1000, 0.01, 400,
     [(1, 16, 64), (0, 1, 128), (0, 2, 128), (0, 4, 128), (0, 6, 128), (0, 8, 128), (0, 10, 128), (0, 12, 128)]),
    ("alpha_dc dt=.01", 2, priors.draw_prior_batch("alpha_dc", 20_000, np.random.default_rng(4)), 1000, 0.01, 400,
     [(1, 16, 64), (0, 2, 128), (0, 4, 128), (0, 8, 128), (0, 10, 128)]),
    ("alpha dt=.001 (fine)", 1, priors.draw_prior_batch("alpha", 20_000, np.random.default_rng(5)), 1000, 1e-3, 4000,
     [(1, 5, 64), (0, 2, 128), (0, 3, 128), (0, 4, 128), (0, 5, 128)]),
]
for name, model, params, n, dt, ms, grid in cases:
    c1, c0 = checksum(model, params, n, dt, ms, 1), checksum(model, params, n, dt, ms, 0)
    print(json.dumps({"case": name, "identical_output_checksums": c1 == c0, "legacy": c1, "tile": c0}), flush=True)
    for variant, thr, tile in grid:
        ms_k, st = run(model, params, n, dt, ms, variant, thr, tile)
        print(json.dumps({"case": name, "variant": "legacy" if variant else "tile", "thr": thr, "tile": st["tile"], "grid": st["grid"],
                          "kernel_ms": round(ms_k, 4), "steps_per_s": st["total_steps"] / (ms_k * 1e-3),
                          "trials_per_s": st["n_trials"] / (ms_k * 1e-3), "steps_per_trial": st["total_steps"] / st["n_trials"]}), flush=True)
sim.close()
