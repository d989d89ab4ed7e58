"""Generate tests/golden/reference_two_channel.npz from the UNMODIFIED reference (build container only):
the retired zoo's two-latent / two-channel simulators under numba with an in-jit seed.
    python tests/golden/make_golden_two_channel.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_two_channel.npz")
CASES = [
    # name, variant, params (the reference's order), n_trials, seed
    ("dc5_typical", "drift_dc5", [1.5, 1.2, 0.5, 0.4, 1.0, 1.0, 0.5, 0.7, -0.4, 0.3, 0.6], 300, 31),
    ("dc5_rejections", "drift_dc5", [-0.5, 0.9, 0.4, 0.3, 2.0, 0.3, 1.5, -1.2, 0.8, 0.9, 0.1], 300, 32),
    ("dc5_timeouts", "drift_dc5", [0.0, 3.0, 0.5, 0.3, 0.2, 0.4, 0.1, 0.5, 0.5, 0.5, 0.5], 100, 33),
    ("dc4_typical", "drift_dc4", [1.5, 1.2, 0.5, 0.4, 1.0, 1.0, 0.5, 0.7, -0.4, 0.3, 0.6], 300, 34),
    ("alpha_dc_typical", "alpha_dc2ch", [2.0, 1.3, 0.55, 0.35, 0.4, 1.0, 0.5, 0.6, -0.3, 1.0, 2.0], 300, 35),
    ("alpha_dc_rejections", "alpha_dc2ch", [0.5, 0.3, 0.45, 0.3, 1.5, 0.4, 1.2, 1.1, 0.9, 0.2, 0.2], 300, 36),
    ("drift_alpha_typical", "drift_alpha", [1.0, 1.3, 0.5, 0.4, 1.0, 1.1, 0.5, 0.6, -0.7, 0.4, 0.8], 300, 37),
    ("alpha_std1_typical", "alpha_std1", [2.0, 1.2, 0.5, 0.35, 0.6, 1.0, 0.7], 300, 38),
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
        store[f"{name}__meta"] = np.array([n, seed, {"drift_dc5": 0, "drift_dc4": 1, "alpha_dc2ch": 2, "drift_alpha": 3, "alpha_std1": 4}[variant]], dtype=np.int64)
        print(f"{name:22s} {variant:12s} shape={out.shape} missing={(out[:, 0] == 0).sum()} cols={out.shape[1]} eeg1 sd={out[:, 1].std():.3f}")
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
