"""Exact first-passage law of the DISCRETE Euler-Maruyama chain -- TEST INFRASTRUCTURE ONLY.

The reference's simulator (basic_ddm_dc.py:95-101) is the Markov chain
    x_{n+1} = x_n + drift*dt + sqrt(dt)*dc*z_n,   absorbed when x <= 0 or x >= bound,
stopped at n = max_steps.  At dt = .01 its first-passage distribution is
visibly different from the continuous-time Wiener law (RTs live on the lattice
n*dt; boundary overshoot shortens RTs), so a KS test against the continuous
WFPT CDF rejects for the wrong reason at large sample sizes.  This module
propagates the sub-density of the chain on a grid (midpoint quadrature of the
Gaussian transition kernel; absorbed mass per step from exact normal tails),
which is exact for the chain up to O(h^2) quadrature error.
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtr


def first_passage_pmf(drift, bound, beta, dc, dt=0.01, max_steps=400, grid=1500):
    """Return (p_upper[n], p_lower[n], p_timeout) for n = 0..max_steps.

    p_upper[n] = P(chain absorbed at the upper boundary exactly at step n) etc.
    Step 0 carries mass only when the start point is already outside (0, bound).
    """
    max_steps = int(max_steps)
    mu = drift * dt
    sd = np.sqrt(dt) * dc
    pu = np.zeros(max_steps + 1)
    pl = np.zeros(max_steps + 1)
    x0 = bound * beta
    if not (x0 > 0):
        pl[0] = 1.0
        return pu, pl, 0.0
    if not (x0 < bound):
        pu[0] = 1.0
        return pu, pl, 0.0
    h = bound / grid
    x = (np.arange(grid) + 0.5) * h
    # step 1 from the point mass
    pu[1] = 1.0 - ndtr((bound - x0 - mu) / sd)
    pl[1] = ndtr((0.0 - x0 - mu) / sd)
    dens = np.exp(-0.5 * ((x - x0 - mu) / sd) ** 2) / (sd * np.sqrt(2 * np.pi))
    # transition kernel K[i, j] = pdf(x_i | from x_j)
    K = np.exp(-0.5 * ((x[:, None] - x[None, :] - mu) / sd) ** 2) / (sd * np.sqrt(2 * np.pi)) * h
    up_from = 1.0 - ndtr((bound - x - mu) / sd)
    lo_from = ndtr((0.0 - x - mu) / sd)
    # renormalise the quadrature so that mass is conserved exactly per source cell
    tot = K.sum(0) + up_from + lo_from
    K /= tot[None, :]
    up_from = up_from / tot
    lo_from = lo_from / tot
    # make step-1 density consistent with its absorbed mass
    alive = 1.0 - pu[1] - pl[1]
    s = dens.sum() * h
    if s > 0:
        dens *= alive / s
    for n in range(2, max_steps + 1):
        m = dens * h
        pu[n] = float(up_from @ m)
        pl[n] = float(lo_from @ m)
        dens = (K @ m) / h
        if dens.sum() * h < 1e-15:
            break
    p_timeout = max(0.0, 1.0 - pu.sum() - pl.sum())
    return pu, pl, p_timeout


def signed_step_cdf(pu, pl, p_timeout):
    """CDF over the signed step count S = choice*n (timeouts mapped to S=0):
    support -max..-1, 0, 1..max.  Returns (support, cdf)."""
    max_steps = len(pu) - 1
    support = np.arange(-max_steps, max_steps + 1)
    pmf = np.concatenate([pl[:0:-1], [p_timeout + pu[0] + pl[0]], pu[1:]])
    return support, np.cumsum(pmf)
