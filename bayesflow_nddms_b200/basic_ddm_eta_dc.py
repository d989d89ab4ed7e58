"""Drop-in for the simulator of the reference's ``retired_models/basic_ddm_eta_dc.py`` (lines 42-120):
the dcDDM with trial-to-trial drift variability, drift_trial ~ N(mu_drift, eta).  This is also the
Euler-Maruyama counterpart of the (mu, eta) data generator the JAGS/Stan scripts obtain from
``simulratcliff(Nu=..., Eta=...)`` (alpha_not_scaled.py:95-97).

``simulate_trials(params, n_trials) -> (n_trials, 2) float64`` with columns (rt, choice); parameters
``[mu_drift, alpha, beta, ter, eta, dc]``.  A timeout reports choice = 0 (see basic_ddm_dc).
"""
from __future__ import annotations

import numpy as np

from . import _capi, priors
from ._model_common import ModelAPI, configurator, device_configurator  # noqa: F401
from .priors import prior_N, truncnorm_better  # noqa: F401

NUMBA_TIMEOUT_COMPAT = False
RNG = np.random.default_rng(2023)
_api = ModelAPI(_capi.MODEL_ETA, "eta")
num_params = 6


def _flags():
    return _capi.FLAG_TIMEOUT_CHOICE_ONE if NUMBA_TIMEOUT_COMPAT else 0


def draw_prior():
    """basic_ddm_eta_dc.py:54-75 -> (6,) [mu_drift, alpha, beta, ter, eta, dc]."""
    return priors.draw_prior_batch("eta", 1, RNG)[0]


def batch_draw_prior(batch_size, *args, **kwargs):
    return priors.draw_prior_batch("eta", batch_size, RNG)


def diffusion_trial(mu_drift, alpha, beta, ter, eta, dc, dt=.01, max_steps=400., simulator=None):
    """basic_ddm_eta_dc.py:80-107 -> (rt, choice)."""
    out = _api.batch_simulate_trials(np.array([[mu_drift, alpha, beta, ter, eta, dc]]), 1, simulator, dt=dt,
                                     max_steps=max_steps, flags=_flags())
    return float(out[0, 0, 0]), int(out[0, 0, 1])


def simulate_trials(params, n_trials, simulator=None):
    """basic_ddm_eta_dc.py:109-120 -> (n_trials, 2) float64."""
    return _api.simulate_trials(params, n_trials, simulator, flags=_flags())


def batch_simulate_trials(params, n_trials, simulator=None, **kw):
    kw.setdefault("flags", _flags())
    return _api.batch_simulate_trials(params, n_trials, simulator, **kw)


def batch_simulate_trials_device(params, n_trials, simulator=None, **kw):
    kw.setdefault("flags", _flags())
    return _api.batch_simulate_trials_device(params, n_trials, simulator, **kw)


def generative_model(batch_size, simulator=None, device=False, device_prior=False):
    return _api.generative_model(batch_size, batch_draw_prior, prior_N, simulator, device, device_prior)
