"""Generate tests/golden/reference_evidence.npz from the UNMODIFIED reference (build container only).

Evidence-path simulators of the retired model zoo, run verbatim under numba
with an in-jit seed:  python tests/golden/make_golden_evidence.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_evidence.npz")

CASES = [
    # name, variant, params, n_trials, seed
    ("evidence_default", "evidence", [3.0, 1.0, 0.5, 0.4, 1.0, 1.0], 40, 21),
    ("evidence_slow", "evidence", [0.2, 2.5, 0.45, 0.3, 0.6, 0.3], 30, 22),        # paths longer than 200 steps, some timeouts
    ("evidence_fast", "evidence", [4.0, 0.5, 0.5, 0.2, 1.5, 2.0], 30, 23),         # paths much shorter than 200 steps
    ("evidence2_default", "evidence2", [3.0, 1.0, 0.5, 0.4, 1.0, 1.0], 40, 24),
    ("evidence2_slow", "evidence2", [-0.3, 2.0, 0.55, 0.35, 0.7, 0.5], 25, 25),
    ("evidence_nn2_default", "evidence_no_noise2", [3.0, 1.0, 0.5, 0.4, 1.0], 30, 26),
    ("evidence_nn2_slow", "evidence_no_noise2", [0.5, 2.2, 0.5, 0.3, 0.8], 20, 27),
]


def main():
    if not rl.available():
        raise SystemExit("reference tree not found; run in the build container")
    store = {}
    for name, variant, params, n, seed in CASES:
        p = np.asarray(params, dtype=np.float64)
        out = np.asarray(rl.seeded_call(variant, seed, p, n), dtype=np.float64)
        store[f"{name}__out"] = out
        store[f"{name}__params"] = p
        store[f"{name}__meta"] = np.array([n, seed, {"evidence": 1, "evidence2": 2, "evidence_no_noise2": 1}[variant],
                                           out.shape[1] - 2], dtype=np.int64)
        print(f"{name:22s} {variant:20s} shape={out.shape} mean rt={out[:, 0].mean():.3f} timeouts={(out[:, 1] == 0).sum()}")
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
