"""Sweep the persistent kernel's tuning knobs on the C5 workload (device-resident timing)."""
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg  # noqa: E402
from bayesflow_nddms_b200 import priors  # noqa: E402

D = int(os.environ.get("TUNE_DATASETS", "100000"))
sim = pkg.DDMSimulator(0, seed=2023)
params = priors.draw_prior_batch("sweep", D, np.random.default_rng(2023))
sim._check(sim._lib.ddm_upload_params(sim._ctx, 0, params.ctypes.data_as(pkg._capi._dp), D, 5))


def run(thr, bps, tile, reps=3):
    sim.set_tuning(thr, bps, tile)
    best = None
    for _ in range(reps):
        sim._check(sim._lib.ddm_run(sim._ctx, 1000, 1e-3, 4000, 2023, 0, 32, 2))
        st = sim.last_stats()
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    return best


rows = []
grid = list(itertools.product([2, 4, 6, 8, 12, 16], [0, 4, 5], [32, 64, 256]))
if len(sys.argv) > 1:
    grid = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for thr, bps, tile in grid:
    st = run(thr, bps, tile)
    sps = st["total_steps"] / (st["kernel_ms"] * 1e-3)
    rows.append(dict(thr=thr, bps=bps, tile=tile, ms=st["kernel_ms"], steps_per_s=sps, grid=st["grid"]))
    print(f"thr={thr:2d} bps={bps} tile={tile:3d} grid={st['grid']:4d} ms={st['kernel_ms']:8.3f} steps/s={sps:.4e}", flush=True)
best = max(rows, key=lambda r: r["steps_per_s"])
print("BEST", json.dumps(best))
