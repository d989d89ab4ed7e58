"""Exact law of the CUDA kernels' discrete Box-Muller map -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

bayesflow_nddms_b200/csrc/ddm_rng.cuh turns two 21-bit fields (m, k) of a Philox block into
    u = (2m + 1) / 2^22,  t = k / 2^21 - 1/2,  r = sqrt(-2 ln u),  z_even = r cos(2 pi t),  z_odd = r sin(2 pi t).
The radius takes 2^21 values and is capped at sqrt(2 * 22 * ln 2) = 5.5226; the reference draws full-range fp64
normals (numba's MT19937 + polar method).  This module computes the law of |z| under the map exactly in the radius
(the 2^21 angles are treated as continuous: their lattice is 3e-6 rad fine) and its distance from N(0, 1), so that
the deviation is a number in DESIGN.md and the device histogram has something exact to be tested against.
"""
from __future__ import annotations

import math

import numpy as np

N_FIELD = 1 << 21
Z_CAP = math.sqrt(2.0 * 22.0 * math.log(2.0))


def radii() -> np.ndarray:
    m = np.arange(N_FIELD, dtype=np.float64)
    return np.sqrt(-2.0 * np.log((2.0 * m + 1.0) / 4194304.0))


def abs_cdf(x, r=None) -> np.ndarray:
    """P(|z| <= x) under the discrete map, for an array of x >= 0."""
    r = radii() if r is None else r
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    out = np.empty(x.size)
    for i, xi in enumerate(x):
        q = np.minimum(xi / r, 1.0)
        out[i] = 1.0 - (2.0 / math.pi) * float(np.mean(np.arccos(q)))
    return out


def normal_abs_cdf(x) -> np.ndarray:
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    return np.array([math.erf(v / math.sqrt(2.0)) for v in x])


def bin_probabilities(edges, r=None):
    """(p_map, p_normal) per bin of |z| for the given edges, plus the mass beyond the last edge."""
    cm, cn = abs_cdf(edges, r), normal_abs_cdf(edges)
    return np.diff(cm), np.diff(cn), 1.0 - cm[-1], 1.0 - cn[-1]


def total_variation(step: float = 0.0025, r=None) -> float:
    """TV distance between the law of |z| (hence of z, both being symmetric) under the map and under N(0, 1),
    on a grid of `step` (a lower bound that converges from below as the grid is refined)."""
    edges = np.arange(0.0, 6.5 + step, step)
    pm, pn, tm, tn = bin_probabilities(edges, r)
    return 0.5 * (float(np.abs(pm - pn).sum()) + abs(tm - tn))


def lost_tail_mass() -> float:
    """P(|Z| > Z_CAP) for a true standard normal: mass the map cannot produce."""
    return math.erfc(Z_CAP / math.sqrt(2.0))
