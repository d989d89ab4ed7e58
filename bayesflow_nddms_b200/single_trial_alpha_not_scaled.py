"""Drop-in for the hot path of the reference's ``single_trial_alpha_not_scaled.py``.

Covers the model itself (lines 66-191: per-trial boundary ~ N(mu_alpha, std_alpha) redrawn
until positive, ext-data = N(bound_trial, sigma1), output (signed choicert, extdata1)) and
its misspecification simulators: ``_alt`` (per-trial diffusion coefficient, :899-974),
``_scale`` (ext-data = N(gamma*bound, sigma1), :1205-1285), ``_scale2`` (gamma = 2,
:1471-1519) and ``_fine`` (dt = .001, max_steps = 4000, :1710-1722).
"""
from __future__ import annotations

import numpy as np

from . import _capi, priors
from ._model_common import ModelAPI, bayesflow_generative_model, configurator, device_configurator  # noqa: F401
from .priors import prior_N, truncnorm_better  # noqa: F401

RNG = np.random.default_rng(2023)
num_params = 7

_api = ModelAPI(_capi.MODEL_ALPHA, "alpha")
_api_alt = ModelAPI(_capi.MODEL_ALPHA_DC, "alpha_dc")
_api_scale = ModelAPI(_capi.MODEL_ALPHA_SCALE, "alpha_scale")
_api_scale2 = ModelAPI(_capi.MODEL_ALPHA_SCALE2, "alpha_scale2")
_api_fine = ModelAPI(_capi.MODEL_ALPHA, "alpha", dt=.001, max_steps=4000.)


def _trial(api, params, dt, max_steps, simulator):
    out = api.batch_simulate_trials(np.asarray(params, dtype=np.float64)[None, :], 1, simulator, dt=dt,
                                    max_steps=max_steps)
    return float(out[0, 0, 0]), float(out[0, 0, 1])


# ---- the model (lines 66-191) ---------------------------------------------------------------
def draw_prior():
    """:78-102 -> (7,) [drift, mu_alpha, beta, ter, std_alpha, dc, sigma1]."""
    return priors.draw_prior_batch("alpha", 1, RNG)[0]


def batch_draw_prior(batch_size, *args, **kwargs):
    return priors.draw_prior_batch("alpha", batch_size, RNG)


def diffusion_trial(drift, mu_alpha, beta, ter, std_alpha, dc, sigma1, dt=.01, max_steps=400., simulator=None):
    """:107-142 -> (choicert, extdata1)."""
    return _trial(_api, [drift, mu_alpha, beta, ter, std_alpha, dc, sigma1], dt, max_steps, simulator)


def simulate_trials(params, n_trials, simulator=None):
    """:144-155 -> (n_trials, 2) float64: col0 signed choicert (0 = missing), col1 extdata1."""
    return _api.simulate_trials(params, n_trials, simulator)


def batch_simulate_trials(params, n_trials, simulator=None, **kw):
    return _api.batch_simulate_trials(params, n_trials, simulator, **kw)


def batch_simulate_histogram(params, n_trials, simulator=None, **kw):
    """(B, P) host parameters -> RT histogram by boundary of the B x n_trials simulated trials, reduced on the GPU."""
    return _api.batch_simulate_histogram(params, n_trials, simulator, **kw)


def batch_simulate_trials_device(params, n_trials, simulator=None, **kw):
    return _api.batch_simulate_trials_device(params, n_trials, simulator, **kw)


def generative_model(batch_size, simulator=None, device=False, device_prior=False):
    return _api.generative_model(batch_size, batch_draw_prior, prior_N, simulator, device, device_prior)


# ---- per-trial diffusion coefficient (:899-974) ------------------------------------------------
def draw_prior_alt():
    """:899-923 -> (7,) [drift, alpha, beta, ter, std_dc, mu_dc, sigma1]."""
    return priors.draw_prior_batch("alpha_dc", 1, RNG)[0]


def batch_draw_prior_alt(batch_size, *args, **kwargs):
    return priors.draw_prior_batch("alpha_dc", batch_size, RNG)


def diffusion_trial_alt(drift, alpha, beta, ter, std_dc, mu_dc, sigma1, dt=.01, max_steps=400., simulator=None):
    return _trial(_api_alt, [drift, alpha, beta, ter, std_dc, mu_dc, sigma1], dt, max_steps, simulator)


def simulate_trials_alt(params, n_trials, simulator=None):
    return _api_alt.simulate_trials(params, n_trials, simulator)


def batch_simulate_trials_alt(params, n_trials, simulator=None, **kw):
    return _api_alt.batch_simulate_trials(params, n_trials, simulator, **kw)


# ---- ext-data scaled by gamma (:1205-1285) and by 2 (:1471-1519) ----------------------------------
def draw_prior_scale():
    """:1205-1232 -> (8,) [..., sigma1, gamma]."""
    return priors.draw_prior_batch("alpha_scale", 1, RNG)[0]


def batch_draw_prior_scale(batch_size, *args, **kwargs):
    return priors.draw_prior_batch("alpha_scale", batch_size, RNG)


def diffusion_trial_scale(drift, mu_alpha, beta, ter, std_alpha, dc, sigma1, gamma, dt=.01, max_steps=400.,
                          simulator=None):
    return _trial(_api_scale, [drift, mu_alpha, beta, ter, std_alpha, dc, sigma1, gamma], dt, max_steps, simulator)


def simulate_trials_scale(params, n_trials, simulator=None):
    return _api_scale.simulate_trials(params, n_trials, simulator)


def batch_simulate_trials_scale(params, n_trials, simulator=None, **kw):
    return _api_scale.batch_simulate_trials(params, n_trials, simulator, **kw)


def diffusion_trial_scale2(drift, mu_alpha, beta, ter, std_alpha, dc, sigma1, dt=.01, max_steps=400., simulator=None):
    return _trial(_api_scale2, [drift, mu_alpha, beta, ter, std_alpha, dc, sigma1], dt, max_steps, simulator)


def simulate_trials_scale2(params, n_trials, simulator=None):
    return _api_scale2.simulate_trials(params, n_trials, simulator)


def batch_simulate_trials_scale2(params, n_trials, simulator=None, **kw):
    return _api_scale2.batch_simulate_trials(params, n_trials, simulator, **kw)


# ---- finer time step (:1710-1722) -----------------------------------------------------------------
def simulate_trials_fine(params, n_trials, simulator=None):
    """The model simulated with dt = .001 and max_steps = 4000."""
    return _api_fine.simulate_trials(params, n_trials, simulator)


def batch_simulate_trials_fine(params, n_trials, simulator=None, **kw):
    return _api_fine.batch_simulate_trials(params, n_trials, simulator, **kw)


def make_bayesflow_generative_model(batched=True):
    return bayesflow_generative_model(draw_prior, prior_N, simulate_trials,
                                      batch_simulate_trials if batched else None,
                                      batch_draw_prior if batched else None)
