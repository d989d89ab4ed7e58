"""Drop-in for the hot path of the reference's ``basic_ddm_dc.py`` (lines 50-160).

Same names, argument meaning and output layout as the reference:
``prior_N``, ``truncnorm_better``, ``draw_prior() -> (5,)``, ``diffusion_trial``,
``simulate_trials(params, n_trials) -> (n_trials, 2) float64`` with columns
(rt = n*dt + tau, choice in {+1, -1}), ``configurator``.  Additions: ``batch_*`` variants
(BayesFlow's batched mode: one kernel launch per batch) and device-resident output.

Timeouts: the reference leaves ``choice`` unbound at basic_ddm_dc.py:110-112; under numba the
trial reports choice = 1.  Here a timeout reports choice = 0 (the author's stated "missing
response"); set ``NUMBA_TIMEOUT_COMPAT = True`` to reproduce the numba artefact.
"""
from __future__ import annotations

import numpy as np

from . import _capi, priors
from ._model_common import ModelAPI, bayesflow_generative_model, configurator, device_configurator  # noqa: F401
from .priors import prior_N, truncnorm_better  # noqa: F401

NUMBA_TIMEOUT_COMPAT = False
RNG = np.random.default_rng(2023)
_api = ModelAPI(_capi.MODEL_BASIC, "basic")
num_params = 5


def _flags():
    return _capi.FLAG_TIMEOUT_CHOICE_ONE if NUMBA_TIMEOUT_COMPAT else 0


def draw_prior():
    """basic_ddm_dc.py:62-80 -> (5,) float64 [drift, alpha, beta, ter, dc]."""
    return priors.draw_prior_batch("basic", 1, RNG)[0]


def batch_draw_prior(batch_size, *args, **kwargs):
    """``Prior(batch_prior_fun=...)`` contract: (batch_size, 5) float64."""
    return priors.draw_prior_batch("basic", batch_size, RNG)


def diffusion_trial(drift, boundary, beta, tau, dc, dt=.01, max_steps=400., simulator=None):
    """basic_ddm_dc.py:85-112 -> (rt, choice).  One trial = one launch; use simulate_trials."""
    out = _api.batch_simulate_trials(np.array([[drift, boundary, beta, tau, dc]]), 1, simulator, dt=dt,
                                     max_steps=max_steps, flags=_flags())
    return float(out[0, 0, 0]), int(out[0, 0, 1])


def simulate_trials(params, n_trials, simulator=None):
    """basic_ddm_dc.py:114-125 -> (n_trials, 2) float64."""
    return _api.simulate_trials(params, n_trials, simulator, flags=_flags())


def batch_simulate_trials(params, n_trials, simulator=None, **kw):
    """(B, 5), int -> (B, n_trials, 2) float64, one launch."""
    kw.setdefault("flags", _flags())
    return _api.batch_simulate_trials(params, n_trials, simulator, **kw)


def batch_simulate_histogram(params, n_trials, simulator=None, **kw):
    """(B, P) host parameters -> RT histogram by boundary of the B x n_trials simulated trials, reduced on the GPU."""
    return _api.batch_simulate_histogram(params, n_trials, simulator, **kw)


def batch_simulate_trials_device(params, n_trials, simulator=None, **kw):
    kw.setdefault("flags", _flags())
    return _api.batch_simulate_trials_device(params, n_trials, simulator, **kw)


def generative_model(batch_size, simulator=None, device=False, device_prior=False):
    """Result dict of ``GenerativeModel(prior, simulator)(batch_size)`` (basic_ddm_dc.py:134)."""
    return _api.generative_model(batch_size, batch_draw_prior, prior_N, simulator, device, device_prior)


def make_bayesflow_generative_model(batched=True):
    return bayesflow_generative_model(draw_prior, prior_N, simulate_trials,
                                      batch_simulate_trials if batched else None,
                                      batch_draw_prior if batched else None)
