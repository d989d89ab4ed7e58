"""numba restatement of the reference's trial loop -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The reference executes its simulator under numba (``@njit``, basic_ddm_dc.py:85,114), drawing
normals from numba's per-thread MT19937 with ``np.random.normal()``.  ``/root/reference`` does not
travel to the GPU box, so bench.py's CPU baseline cannot exec the reference there; the C port in
ddm_oracle.c is the measured baseline (``kind: "port"``), and this module adds the same loop
compiled by the reference's own engine, so that the report can state both.  It follows
basic_ddm_dc.py:85-125 (float step counter, strict inequalities, one normal per step,
rt = n*dt + tau, a timeout reported as choice 0) over a (B, 5) parameter matrix.
"""
from __future__ import annotations

import numpy as np

try:
    from numba import njit, prange

    HAVE_NUMBA = True
except Exception:  # pragma: no cover
    HAVE_NUMBA = False


if HAVE_NUMBA:

    @njit(cache=False)
    def _one_dataset(p, n_trials, dt, max_steps, out):
        drift, bound, beta, tau, dc = p[0], p[1], p[2], p[3], p[4]
        total = 0.0
        for i in range(n_trials):
            n = 0.0
            ev = bound * beta
            while ev > 0 and ev < bound and n < max_steps:
                ev += drift * dt + np.sqrt(dt) * dc * np.random.normal()
                n += 1.0
            out[i, 0] = n * dt + tau
            out[i, 1] = 1.0 if ev >= bound else (-1.0 if ev <= 0 else 0.0)
            total += n
        return total

    @njit(cache=False)
    def simulate_serial(params, n_trials, dt, max_steps, out):
        """One thread, datasets in order: how the reference runs under BayesFlow."""
        total = 0.0
        for b in range(params.shape[0]):
            total += _one_dataset(params[b], n_trials, dt, max_steps, out[b])
        return total

    @njit(parallel=True, cache=False)
    def simulate_parallel(params, n_trials, dt, max_steps, out):
        """numba threads over datasets (courtesy upper bound; the reference has no parallelism)."""
        totals = np.zeros(params.shape[0])
        for b in prange(params.shape[0]):
            totals[b] = _one_dataset(params[b], n_trials, dt, max_steps, out[b])
        return totals.sum()


def time_numba(params, n_trials, dt, max_steps, parallel=False):
    """(steps, seconds) of one pass, compile excluded."""
    import time

    if not HAVE_NUMBA:
        raise RuntimeError("numba is not installed")
    fn = simulate_parallel if parallel else simulate_serial
    warm = np.empty((2, 4, 2))
    fn(params[:2].copy(), 4, dt, float(max_steps), warm)  # JIT compile
    out = np.empty((params.shape[0], n_trials, 2))
    t0 = time.perf_counter()
    steps = fn(params, n_trials, dt, float(max_steps), out)
    return float(steps), time.perf_counter() - t0
