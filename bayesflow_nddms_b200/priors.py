"""Prior draws of the dcDDM family -- same distributions and parameter order as the reference.

Reference: basic_ddm_dc.py:50-80 (prior_N, truncnorm_better, draw_prior),
single_trial_alpha_not_scaled.py:66-102, :899-923 (draw_prior_alt), :1205-1232
(draw_prior_scale), imputation_from_stahl_not_scaled.py:165-174 (per-participant draws).

The reference draws one parameter vector per call through scipy's ``truncnorm.rvs``
(about 0.33 ms per call).  Here truncated normals are sampled by inverse
CDF on a ``numpy.random.Generator``, vectorised over the batch; ``draw_prior()`` keeps the
reference's zero-argument signature and (P,) float64 return.
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtr, ndtri

# column names, in the reference's order
PARAM_NAMES = {
    "basic": ("drift", "alpha", "beta", "ter", "dc"),
    "alpha": ("drift", "mu_alpha", "beta", "ter", "std_alpha", "dc", "sigma1"),
    "alpha_dc": ("drift", "alpha", "beta", "ter", "std_dc", "mu_dc", "sigma1"),
    "alpha_scale": ("drift", "mu_alpha", "beta", "ter", "std_alpha", "dc", "sigma1", "gamma"),
    "alpha_scale2": ("drift", "mu_alpha", "beta", "ter", "std_alpha", "dc", "sigma1"),
    "stahl": ("drift", "beta", "ter", "dc"),
    "eta": ("mu_drift", "alpha", "beta", "ter", "eta", "dc"),
    # throughput sweep C5: the basic prior with tau = 0
    "sweep": ("drift", "alpha", "beta", "ter", "dc"),
}


def prior_N(n_min=60, n_max=300):
    """Number of trials shared by a batch, U{n_min..n_max} (basic_ddm_dc.py:50-52)."""
    return np.random.randint(n_min, n_max + 1)


def truncnorm_rvs(rng, mean, sd, low, upp, size):
    """Normal(mean, sd) truncated to [low, upp], inverse-CDF sampling."""
    a, b = ndtr((low - mean) / sd), ndtr((upp - mean) / sd)
    u = rng.uniform(0.0, 1.0, size)
    x = mean + sd * ndtri(a + u * (b - a))
    return np.clip(x, low, upp)


def truncnorm_better(mean=0, sd=1, low=-10, upp=10, size=1, rng=None):
    """Same signature and return shape as the reference helper (basic_ddm_dc.py:55-57)."""
    return truncnorm_rvs(np.random.default_rng() if rng is None else rng, mean, sd, low, upp, size)


def draw_prior_batch(model: str, batch_size: int, rng) -> np.ndarray:
    """(batch_size, P) float64 prior draws for ``model`` (a key of PARAM_NAMES)."""
    B = int(batch_size)
    drift = rng.normal(0.0, 2.0, B)
    alpha = truncnorm_rvs(rng, 1.0, 0.5, 0.0, 10.0, B)
    beta = rng.beta(2.0, 2.0, B)
    ter = truncnorm_rvs(rng, 0.5, 0.25, 0.0, 1.5, B)
    if model == "basic":
        dc = truncnorm_rvs(rng, 1.0, 0.5, 0.0, 10.0, B)
        cols = (drift, alpha, beta, ter, dc)
    elif model == "sweep":
        dc = truncnorm_rvs(rng, 1.0, 0.5, 0.0, 10.0, B)
        cols = (drift, alpha, beta, np.zeros(B), dc)
    elif model in ("alpha", "alpha_dc", "alpha_scale", "alpha_scale2"):
        std = truncnorm_rvs(rng, 1.0, 0.5, 0.0, 3.0, B)     # std_alpha | std_dc
        dc = truncnorm_rvs(rng, 1.0, 0.5, 0.0, 10.0, B)     # dc | mu_dc
        sigma1 = rng.uniform(0.0, 5.0, B)
        cols = (drift, alpha, beta, ter, std, dc, sigma1)
        if model == "alpha_scale":
            cols = cols + (rng.uniform(0.0, 2.0, B),)
    elif model == "eta":
        # retired_models/basic_ddm_eta_dc.py:54-75
        eta = truncnorm_rvs(rng, 1.0, 0.5, 0.0, 3.0, B)
        dc = truncnorm_rvs(rng, 1.0, 0.5, 0.0, 10.0, B)
        cols = (drift, alpha, beta, ter, eta, dc)
    elif model == "stahl":
        # imputation_from_stahl_not_scaled.py:165-174
        cols = (rng.normal(3.0, 1.0, B), rng.beta(25.0, 25.0, B), truncnorm_rvs(rng, 0.4, 0.1, 0.0, 1.5, B),
                truncnorm_rvs(rng, 1.0, 0.25, 0.0, 10.0, B))
    else:
        raise ValueError(f"unknown prior {model!r}")
    return np.stack(cols, axis=-1).astype(np.float64)
