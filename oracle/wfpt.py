"""Wiener first-passage-time density / CDF (Navarro & Fuss 2009) -- TEST INFRASTRUCTURE ONLY.

The reference contains no WFPT code of its own (``logwienerpdf`` was deleted,
pyhddmjagsutils.py:25,32); its analytic comparators call third-party
likelihoods -- JAGS ``dwiener`` (jags-wiener module, version unpinned,
basic_ddm_dc_pyjags.py:133) and Stan ``wiener_lpdf`` (pystan 2.19,
stancode/basic_ddm_dc_test.stan:14-27) -- neither of which is installed.
So this file restates the published algorithm (Navarro & Fuss 2009, J. Math.
Psych. 53: small-/large-time series for the unit-diffusion density) and the
reference's dc-scaling identity from those call sites:

    dwiener(alpha/varsigma, ndt, beta, delta/varsigma)     basic_ddm_dc_pyjags.py:133
    upper boundary via (1 - beta, -delta)                  stancode/basic_ddm_dc_test.stan:17-25

PARITY UNPINNED at this boundary: no reference test pins the density.  It is
cross-checked in tests/ against (i) numerical integration to the closed-form
absorption probability, (ii) the exact discrete-chain oracle at small dt,
(iii) the reference's own exact sampler ``simulratcliff`` via golden samples.
"""
from __future__ import annotations

import numpy as np


def prob_lower(v: float, a: float, w: float) -> float:
    """P(absorb at 0) for unit-diffusion Wiener process started at a*w."""
    if abs(v * a) < 1e-10:
        return 1.0 - w
    if v > 0:
        e1 = np.exp(-2.0 * v * a * w)
        e2 = np.exp(-2.0 * v * a)
        return float((e1 - e2) / (1.0 - e2))
    # v < 0: same expression multiplied through by exp(2va) so every exponent is negative
    e1 = np.exp(2.0 * v * a * (1.0 - w))
    e2 = np.exp(2.0 * v * a)
    return float((e1 - 1.0) / (e2 - 1.0))


def _f1_small(u, w, kmax=12):
    k = np.arange(-kmax, kmax + 1)[:, None]
    u = np.asarray(u, dtype=np.float64)[None, :]
    s = ((w + 2 * k) * np.exp(-((w + 2 * k) ** 2) / (2 * u))).sum(0)
    return s / np.sqrt(2 * np.pi * u[0] ** 3)


def _f1_large(u, w, kmax=400):
    k = np.arange(1, kmax + 1)[:, None]
    u = np.asarray(u, dtype=np.float64)[None, :]
    return np.pi * (k * np.exp(-(k ** 2) * np.pi ** 2 * u / 2) * np.sin(k * np.pi * w)).sum(0)


def pdf_lower(t, v, a, w):
    """Defective density of hitting the LOWER boundary at decision time t (unit diffusion)."""
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    out = np.zeros_like(t)
    ok = t > 0
    u = t[ok] / a ** 2
    f1 = np.where(u < 0.35, _f1_small(u, w), _f1_large(u, w))
    out[ok] = np.maximum(f1, 0.0) * np.exp(-v * a * w - v ** 2 * t[ok] / 2) / a ** 2
    return out


def cdf_lower(t, v, a, w, kcap=20000):
    """Defective CDF P(T <= t, lower).  Large-time series integrated term by term:
    f(t) = (pi/a^2) e^{-vaw} sum_k k sin(k pi w) e^{-lam_k t},  lam_k = v^2/2 + k^2 pi^2/(2 a^2)."""
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    out = np.zeros_like(t)
    pl = prob_lower(v, a, w)
    for i, ti in enumerate(t):
        if ti <= 0:
            continue
        K = int(min(kcap, max(20, np.ceil(a * 2.8 / np.sqrt(ti)) + 5)))
        k = np.arange(1, K + 1, dtype=np.float64)
        lam = v ** 2 / 2 + (k * np.pi / a) ** 2 / 2
        tail = (np.pi / a ** 2) * np.exp(-v * a * w) * np.sum(k * np.sin(k * np.pi * w) * np.exp(-lam * ti) / lam)
        out[i] = pl - tail
    return np.clip(out, 0.0, pl)


# --- the reference's parameterisation (drift, boundary, beta, dc) ------------

def ddm_pdf(t, choice, drift, boundary, beta, dc):
    """Defective decision-time density for choice +1 (upper) / -1 (lower)."""
    a, v = boundary / dc, drift / dc
    if choice > 0:
        return pdf_lower(t, -v, a, 1.0 - beta)
    return pdf_lower(t, v, a, beta)


def ddm_cdf(t, choice, drift, boundary, beta, dc):
    a, v = boundary / dc, drift / dc
    if choice > 0:
        return cdf_lower(t, -v, a, 1.0 - beta)
    return cdf_lower(t, v, a, beta)


def ddm_prob_upper(drift, boundary, beta, dc):
    return 1.0 - prob_lower(drift / dc, boundary / dc, beta)


def signed_rt_cdf(x, drift, boundary, beta, dc):
    """CDF of the signed decision time S = choice * T on the real line
    (a proper CDF; the natural one-sample KS target for (rt, choice) data)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    pl = 1.0 - ddm_prob_upper(drift, boundary, beta, dc)
    out = np.empty_like(x)
    neg = x < 0
    # S <= x < 0  <=>  lower and T >= -x
    out[neg] = pl - ddm_cdf(-x[neg], -1, drift, boundary, beta, dc)
    out[~neg] = pl + ddm_cdf(x[~neg], +1, drift, boundary, beta, dc)
    return out
