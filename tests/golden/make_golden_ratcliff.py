"""Generate tests/golden/simulratcliff_samples.npz from the UNMODIFIED reference.

``pyhddmjagsutils.simulratcliff`` (pyhddmjagsutils.py:47-176) is the reference's exact
(rejection) first-passage sampler, used by its JAGS/Stan data generators
(alpha_not_scaled.py:95-97).  Its samples pin oracle/wfpt.py (the analytic density has no
reference implementation in-tree) and give the GPU simulator an independent target.
Run in the build container only:  python tests/golden/make_golden_ratcliff.py
"""
import importlib
import os
import sys
import types

import numpy as np

REF = os.environ.get("DDM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simulratcliff_samples.npz")

CASES = {
    # name: (Alpha, Tau, Nu, Beta, Eta, Varsigma), N, seed
    "fast_unbiased": (dict(Alpha=1.5, Tau=.4, Nu=3.0, Beta=.5, Eta=0, Varsigma=1.0), 4000, 11),
    "slow_biased": (dict(Alpha=1.2, Tau=.35, Nu=-1.0, Beta=.4, Eta=0, Varsigma=1.2), 4000, 12),
    "dc_scaled_pair_a": (dict(Alpha=1.2, Tau=.35, Nu=1.5, Beta=.5, Eta=0, Varsigma=1.0), 3000, 13),
    "dc_scaled_pair_b": (dict(Alpha=2.4, Tau=.35, Nu=3.0, Beta=.5, Eta=0, Varsigma=2.0), 3000, 14),
    # participant-17 override of alpha_not_scaled.py:83-88 (drift variability Eta = 1)
    "participant17_eta": (dict(Alpha=1.2, Tau=.4, Nu=3.5, Beta=.5, Eta=1.0, Varsigma=1.2), 4000, 15),
}


def main():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    phju = importlib.import_module("pyhddmjagsutils")
    store = {}
    for name, (kw, n, seed) in CASES.items():
        np.random.seed(seed)
        y = np.real(phju.simulratcliff(N=n, **kw)).astype(np.float64)
        store[f"{name}__y"] = y
        store[f"{name}__params"] = np.array([kw["Alpha"], kw["Tau"], kw["Nu"], kw["Beta"], kw["Eta"], kw["Varsigma"]])
        print(f"{name:20s} n={n} P(upper)={np.mean(y > 0):.3f} mean|rt|={np.mean(np.abs(y)):.3f}")
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
