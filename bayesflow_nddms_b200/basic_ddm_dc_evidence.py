"""Drop-in for the simulators of the reference's evidence-path models (retired model zoo):

  retired_models/basic_ddm_dc_evidence.py:61-151         200 observed samples, noise sigma1, per-trial z-score
  retired_models/basic_ddm_dc_evidence2.py:57-150        same, standardised with the dataset's path-mean statistics
  retired_models/basic_ddm_dc_evidence_no_noise2.py:59-147   400 samples, fixed noise .001, per-trial z-score

``simulate_trials(params, n_trials) -> (n_trials, 2 + n_obs) float64``: rt, choice, observed path
(dt = .001, max_time = 4 s as in the reference's keyword defaults).
"""
from __future__ import annotations

import numpy as np

from . import priors
from ._model_common import configurator, device_configurator  # noqa: F401
from .priors import prior_N, truncnorm_better  # noqa: F401
from .simulator import default_simulator

RNG = np.random.default_rng(2023)
num_params = 6


def draw_prior():
    """basic_ddm_dc_evidence.py:61-82 -> (6,) [drift, alpha, beta, ter, dc, sigma1]."""
    return batch_draw_prior(1)[0]


def batch_draw_prior(batch_size, *args, **kwargs):
    p = priors.draw_prior_batch("basic", batch_size, RNG)
    return np.concatenate([p, RNG.uniform(0.0, 5.0, (int(batch_size), 1))], axis=1)


def draw_prior_no_noise():
    """basic_ddm_dc_evidence_no_noise2.py:59-77 -> (5,) [drift, alpha, beta, ter, dc]."""
    return priors.draw_prior_batch("basic", 1, RNG)[0]


def _run(params, n_trials, n_obs, standardize, simulator, dt, max_time, **kw):
    sim = simulator if simulator is not None else default_simulator()
    params = np.asarray(params, dtype=np.float64)
    single = params.ndim == 1
    p = params[None, :] if single else params
    if p.shape[1] == 5:  # the no-noise variants: fixed .001 "computational" noise (…no_noise2.py:118-120)
        p = np.concatenate([p, np.full((p.shape[0], 1), 0.001)], axis=1)
    out = sim.simulate_evidence(p, n_trials, n_obs, standardize, dt, int(round(max_time / dt)), **kw)
    return out[0] if single and not kw.get("device") else out


def simulate_trials(params, n_trials, simulator=None, dt=.001, max_time=4., **kw):
    """basic_ddm_dc_evidence.py:139-151 -> (n_trials, 202); (B, 6) params give (B, n_trials, 202)."""
    return _run(params, n_trials, int(.2 / dt), 1, simulator, dt, max_time, **kw)


batch_simulate_trials = simulate_trials


def simulate_trials_evidence2(params, n_trials, simulator=None, dt=.001, max_time=4., **kw):
    """basic_ddm_dc_evidence2.py:132-150: paths standardised with mean/std of the dataset's per-trial path means."""
    return _run(params, n_trials, int(.2 / dt), 2, simulator, dt, max_time, **kw)


def simulate_trials_no_noise2(params, n_trials, simulator=None, dt=.001, max_time=4., **kw):
    """basic_ddm_dc_evidence_no_noise2.py:135-147 -> (n_trials, 402), 5 parameters."""
    return _run(params, n_trials, int(.4 / dt), 1, simulator, dt, max_time, **kw)


def diffusion_trial(drift=3, boundary=1, beta=.5, tau=.4, dc=1, sigma1=1, dt=.001, max_time=4., simulator=None):
    """basic_ddm_dc_evidence.py:87-135 -> (rt, choice, obs_path[200])."""
    row = simulate_trials([drift, boundary, beta, tau, dc, sigma1], 1, simulator, dt, max_time)[0]
    return float(row[0]), int(row[1]), row[2:].copy()


def generative_model(batch_size, simulator=None, device=False):
    prior_draws = batch_draw_prior(batch_size)
    n = int(prior_N())
    data = simulate_trials(prior_draws, n, simulator, device=device) if device else simulate_trials(prior_draws, n, simulator)
    return {'prior_draws': prior_draws, 'sim_data': data, 'sim_non_batchable_context': n}
