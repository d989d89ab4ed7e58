"""Host-side mirror of the reference's ``pyhddmjagsutils.simulratcliff`` (pyhddmjagsutils.py:47-176), the exact
first-passage sampler its JAGS / Stan data generators call (alpha_not_scaled.py:95-97), running on the GPU
(``ddm_simulate_exact``).  Only the simulator is mirrored: the module's diagnostics and plotting helpers are out of
scope.

Same signature and return value: ``simulratcliff(N, Alpha, Tau, Nu, Beta, rangeTau, rangeBeta, Eta, Varsigma)`` ->
``(N,) float64`` of response times multiplied by the response (+ upper boundary, - lower).  Streams are Philox,
keyed by the simulator's seed and dataset counter, so runs are reproducible (the reference draws from NumPy's
global state).
"""
from __future__ import annotations

import numpy as np

from .simulator import default_simulator

PARAM_ORDER = ("Alpha", "Tau", "Nu", "Beta", "rangeTau", "rangeBeta", "Eta", "Varsigma")


def simulratcliff(N=100, Alpha=1, Tau=.4, Nu=1, Beta=.5, rangeTau=0, rangeBeta=0, Eta=.3, Varsigma=1, simulator=None, **kw):
    """pyhddmjagsutils.py:47-176 -> (N,) signed response times in seconds."""
    sim = simulator if simulator is not None else default_simulator()
    params = np.array([Alpha, Tau, Nu, Beta, rangeTau, rangeBeta, Eta, Varsigma], dtype=np.float64)
    return sim.simulate_exact(params, int(N), **kw)[0]


def batch_simulratcliff(params, N, simulator=None, **kw):
    """B parameter rows in ``PARAM_ORDER`` -> (B, N): one launch instead of B Python calls
    (alpha_not_scaled.py:95-97 loops over participants and conditions)."""
    sim = simulator if simulator is not None else default_simulator()
    return sim.simulate_exact(params, int(N), **kw)
