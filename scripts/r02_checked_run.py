"""Every kernel family through the bounds-checked build (-DDDM_CHECKED): the in-tree stand-in for compute-sanitizer's
memcheck, which is closed on this GPU pool.  Run as
    DDM_B200_LIB=$PWD/bayesflow_nddms_b200/libddm_b200_checked.so python scripts/r02_checked_run.py
A violated bound is counted in ddm_stats.debug_overruns (and the access skipped); this script asserts zero after every
launch and that the results equal the shipped build's semantics (the one-thread-per-trial kernel, bitwise)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg  # noqa: E402
from bayesflow_nddms_b200 import _capi, priors, two_channel  # noqa: E402
from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as stahl  # noqa: E402

assert "checked" in _capi.library_path(), "set DDM_B200_LIB to the checked build"
sim = pkg.DDMSimulator(0, seed=5)
rng = np.random.default_rng(0)
launches = 0


def clean():
    global launches
    st = sim.last_stats()
    assert st["debug_overruns"] == 0, st
    launches += 1
    return st


MODELS = [(0, "basic"), (1, "alpha"), (2, "alpha_dc"), (3, "alpha_scale"), (4, "alpha_scale2"), (6, "eta")]
for variant in (0, 1):
    sim.set_kernel_variant(variant)
    for thr, bps, tile in ((0, 0, 0), (1, 1, 1), (2, 1, 7), (3, 2, 33), (32, 0, 128), (8, 3, 64), (2, 1, 1000)):
        sim.set_tuning(thr, bps, tile)
        for model, name in MODELS:
            for B, N, kw in ((17, 301, dict(dt=0.01, max_steps=400)), (3, 1000, dict(dt=0.001, max_steps=777)), (1, 1, dict(dt=0.01, max_steps=5)),
                             (65, 33, dict(dt=0.01, max_steps=0))):
                P = priors.draw_prior_batch(name, B, rng)
                a = sim.simulate(model, P, N, seed=3, dataset_offset=2, flags=4, **kw)
                st = clean()
                assert st["used_persistent"] == 1
                b = sim.simulate(model, P, N, seed=3, dataset_offset=2, flags=4 | 8, **kw)
                clean()
                assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (variant, thr, bps, tile, model, B, N)
        raw = np.column_stack([rng.normal(0, 2, 9), 1.0 + rng.random(9), np.full(9, 0.5), np.full(9, 0.3), 0.5 + rng.random(9),
                               0.6 + rng.random(9), np.full(9, 0.3), np.full(9, 0.5), np.full(9, 0.6), np.full(9, 0.2), np.full(9, 0.1)])
        for canon in (two_channel.canonical_drift_dc5(raw), two_channel.canonical_alpha_dc(raw)):
            a = sim.simulate(7, canon, 211, seed=4, dataset_offset=0)
            clean()
            b = sim.simulate(7, canon, 211, seed=4, dataset_offset=0, flags=8)
            clean()
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
sim.set_kernel_variant(0)   # the tile kernel at these sizes, too (the default would take the latency kernel, which has no slots to check)
for thr, bps, tile in ((0, 0, 0), (2, 1, 5), (4, 0, 96)):
    sim.set_tuning(thr, bps, tile)
    n = 20_011
    pp = stahl.draw_participant_params(89, np.random.default_rng(2024))
    group = rng.integers(0, 89, n).astype(np.int32)
    bounds = np.clip(rng.normal(1.0, 0.4, n), 0, None)
    a = sim.simulate_trialwise(group, bounds, pp, seed=3, trial_offset=11)
    clean()
    b = sim.simulate_trialwise(group, bounds, pp, seed=3, trial_offset=11, flags=8)
    clean()
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
sim.set_tuning(0, 0, 0)
sim.set_kernel_variant(-1)
# large launches: many tiles per warp, stragglers, the chunked host path with the compact wire, the histogram call
P = priors.draw_prior_batch("sweep", 6000, rng)
out = sim.simulate(0, P, 1000, 1e-3, 4000, seed=1, dataset_offset=0, flags=2)
clean()
h = sim.simulate_histogram(0, P, 1000, 1e-3, 4000, seed=1, dataset_offset=0, n_bins=100, rt_max=4.0)
st = clean()
assert int(h["upper"].sum() + h["lower"].sum()) + h["missing"] + h["overflow"] == 6000 * 1000
sim.set_pipeline(0, 300_000)   # the same call in chunks (forced; by default from 64 Mi trials on)
h2 = sim.simulate_histogram(0, P, 1000, 1e-3, 4000, seed=1, dataset_offset=0, n_bins=100, rt_max=4.0)
clean()
sim.set_pipeline(-1, -1)
assert np.array_equal(h["upper"], h2["upper"]) and np.array_equal(h["lower"], h2["lower"]) and h["missing"] == h2["missing"]
Pa = priors.draw_prior_batch("alpha", 5000, rng)
sim.simulate(1, Pa, 1000, 0.01, 400, seed=1, dataset_offset=0)
clean()
ev = np.concatenate([priors.draw_prior_batch("basic", 5, rng), np.full((5, 1), 0.5)], axis=1)
for mode in (0, 1, 2):
    sim.simulate_evidence(ev, 64, 200, mode)
    clean()
print(f"checked build: {launches} launches, 0 bounds violations, persistent == generic bitwise everywhere")
sim.close()
