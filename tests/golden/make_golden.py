"""Generate tests/golden/reference_outputs.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

For every hot-path variant it runs the reference's verbatim functions (loaded
by line range, oracle/ref_loader.py) with np.random.seed(seed) called inside
jitted code, and stores inputs + outputs.  The reference has no golden
vectors of its own; these are outputs of the reference
itself and pin the C oracle (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_outputs.npz")

# Fixed-parameter cases the reference itself uses + edge cases.
CASES = [
    # name, variant, params, n_trials, seed
    ("basic_misspec_vector", "basic", [3.0, 1.5, 0.5, 0.4, 1.0], 300, 2023),
    ("basic_neg_drift", "basic", [-1.7, 0.9, 0.3, 0.25, 0.8], 200, 1),
    ("basic_didactic_a", "basic", [1.5, 1.2, 0.5, 0.35, 1.0], 200, 2),
    ("basic_didactic_b", "basic", [3.0, 2.4, 0.5, 0.35, 2.0], 200, 3),
    # slow process: many trials hit max_steps -> numba's undefined-choice artefact (D8)
    ("basic_timeouts", "basic", [0.05, 4.0, 0.5, 0.3, 0.3], 120, 4),
    ("basic_tiny_dc", "basic", [2.0, 1.0, 0.5, 0.1, 1e-3], 50, 5),
    ("basic_one_trial", "basic", [0.3, 1.1, 0.6, 0.5, 1.3], 1, 6),
    ("alpha_misspec_vector", "alpha", [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1], 300, 2023),
    ("alpha_many_rejections", "alpha", [0.5, 0.2, 0.4, 0.3, 2.5, 1.2, 3.0], 300, 8),
    ("alpha_timeouts", "alpha", [0.0, 2.2, 0.5, 0.3, 0.5, 0.45, 1.0], 100, 9),
    ("alpha_dc_vector", "alpha_dc", [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1], 300, 10),
    ("alpha_dc_rejections", "alpha_dc", [-1.0, 1.0, 0.55, 0.2, 2.0, 0.3, 2.0], 200, 11),
    ("alpha_scale_vector", "alpha_scale", [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1, 1.37], 300, 12),
    ("alpha_scale2_vector", "alpha_scale2", [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1], 300, 13),
    ("alpha_fine_vector", "alpha_fine", [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1], 100, 14),
    # retired_models/basic_ddm_eta_dc.py: per-trial drift; participant-17 values of alpha_not_scaled.py:83-88
    ("eta_participant17", "eta", [3.5, 1.2, 0.5, 0.4, 1.0, 1.2], 300, 16),
    ("eta_timeouts", "eta", [0.0, 3.5, 0.5, 0.3, 0.4, 0.35], 100, 17),
]

# imputation_from_stahl_not_scaled.py:120-148 is plain Python on NumPy's global
# (legacy) RandomState, the same MT19937/polar stream.
STAHL = dict(drift=3.2, beta=0.48, ter=0.41, dc=1.05, seed=2024,
             bounds=[1.0, 0.0, 0.37, 2.2, 1.4, 0.05, 3.0, 0.9, 1.1, 0.75, 1.9, 0.6])


def main():
    if not rl.available():
        raise SystemExit("reference tree not found; run in the build container")
    store = {}
    meta = {}
    for name, variant, params, n_trials, seed in CASES:
        p = np.asarray(params, dtype=np.float64)
        out = rl.seeded_call(variant, seed, p, n_trials)
        store[f"{name}__out"] = np.asarray(out, dtype=np.float64)
        store[f"{name}__params"] = p
        meta[name] = dict(variant=variant, n_trials=n_trials, seed=seed)
        print(f"{name:28s} {variant:13s} shape={out.shape} timeouts(col0==0)={int((out[:, 0] == 0).sum())}")

    ns = rl.load("stahl")
    np.random.seed(STAHL["seed"])
    bounds = np.asarray(STAHL["bounds"], dtype=np.float64)
    cr = np.array([ns["diffusion_trial"](STAHL["drift"], b, STAHL["beta"], STAHL["ter"], STAHL["dc"])
                   for b in bounds])
    store["stahl__out"] = cr
    store["stahl__bounds"] = bounds
    store["stahl__params"] = np.array([STAHL["drift"], STAHL["beta"], STAHL["ter"], STAHL["dc"]])
    meta["stahl"] = dict(variant="stahl", n_trials=len(bounds), seed=STAHL["seed"])
    try:
        ns["diffusion_trial"](1.0, -0.1, 0.5, 0.4, 1.0)
        meta["stahl"]["negative_bound_raises"] = False
    except ValueError as e:
        meta["stahl"]["negative_bound_raises"] = True
        meta["stahl"]["negative_bound_message"] = str(e)

    store["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
