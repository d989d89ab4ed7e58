"""Drop-ins for the retired zoo's two-latent / two-channel simulators, on one general kernel
(DDM_MODEL_GENERAL, include/ddm_b200.h):

  retired_models/single_trial_drift_dc5.py:90-154   per-trial drift and diffusion coefficient; two
      standardised external channels, EEG1 ~ N(drift_t + gamma_dc1*dc_t, sigma1), EEG2 ~ N(gamma_dr2*drift_t + dc_t, sigma2)
  retired_models/single_trial_drift_dc4.py:90-146   the same without the standardisation
  retired_models/single_trial_alpha_dc.py:109-176   per-trial boundary and diffusion coefficient, standardised channels
  retired_models/single_trial_drift_alpha.py:96-152 per-trial drift and boundary, two raw channels
  retired_models/single_trial_alpha.py:83-135       per-trial boundary, one standardised channel -> (n_trials, 2)

``simulate_trials_*(params, n_trials) -> (n_trials, 3) float64``: (signed choicert, eeg1, eeg2).
Each variant's parameter vector (the reference's order) is mapped to the kernel's canonical 24
parameters by ``canonical_*``; the standardisation constants are formed here with the reference's
own expressions, so that the fp64 validation kernel reproduces the reference bit for bit.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .simulator import default_simulator

# pre-draw orders (index into the kernel's permutation table of (drift, boundary, dc))
ORDER_DRIFT_BOUND_DC = 0
ORDER_DC_BOUND_DRIFT = 5


def _canon(B):
    c = np.zeros((B, 24), dtype=np.float64)
    c[:, 13] = 1.0   # ext scales
    c[:, 19] = 1.0
    return c


def canonical_drift_dc5(params, standardise=True):
    """[mu_drift, boundary, beta, tau, eta, mu_dc, dc_var, gamma_dc1, gamma_dr2, sigma1, sigma2] -> (B, 24)."""
    p = np.atleast_2d(np.asarray(params, dtype=np.float64))
    mu_drift, boundary, beta, tau, eta, mu_dc, dc_var, g_dc1, g_dr2, s1, s2 = p.T
    c = _canon(p.shape[0])
    c[:, 0], c[:, 1] = mu_drift, eta
    c[:, 2] = boundary
    c[:, 4], c[:, 5] = mu_dc, dc_var
    c[:, 6], c[:, 7] = beta, tau
    c[:, 8], c[:, 10], c[:, 11] = 1.0, g_dc1, s1          # EEG1 = N(1*drift_t + gamma_dc1*dc_t, sigma1)
    c[:, 14], c[:, 16], c[:, 17] = g_dr2, 1.0, s2         # EEG2 = N(gamma_dr2*drift_t + 1*dc_t, sigma2)
    if standardise:                                        # single_trial_drift_dc5.py:122-131
        c[:, 12] = 1 * mu_drift + g_dc1 * mu_dc
        c[:, 13] = np.sqrt(eta ** 2 + (g_dc1 ** 2 * dc_var ** 2) + s1 ** 2)
        c[:, 18] = g_dr2 * mu_drift + mu_dc
        c[:, 19] = np.sqrt((g_dr2 ** 2 * eta ** 2) + dc_var ** 2 + s2 ** 2)
    c[:, 20], c[:, 21], c[:, 22] = ORDER_DRIFT_BOUND_DC, 2, 1
    return c


def canonical_alpha_dc(params):
    """[drift, mu_alpha, beta, tau, var_alpha, mu_dc, var_dc, gamma_dc1, gamma_bd2, sigma1, sigma2] -> (B, 24)."""
    p = np.atleast_2d(np.asarray(params, dtype=np.float64))
    drift, mu_alpha, beta, tau, var_alpha, mu_dc, var_dc, g_dc1, g_bd2, s1, s2 = p.T
    c = _canon(p.shape[0])
    c[:, 0] = drift
    c[:, 2], c[:, 3] = mu_alpha, var_alpha
    c[:, 4], c[:, 5] = mu_dc, var_dc
    c[:, 6], c[:, 7] = beta, tau
    c[:, 9], c[:, 10], c[:, 11] = 1.0, g_dc1, s1          # EEG1 = N(1*bound_t + gamma_dc1*dc_t, sigma1)
    c[:, 15], c[:, 16], c[:, 17] = g_bd2, 1.0, s2         # EEG2 = N(gamma_bd2*bound_t + 1*dc_t, sigma2)
    c[:, 12] = 1 * mu_alpha + g_dc1 * mu_dc               # single_trial_alpha_dc.py:144-154
    c[:, 13] = np.sqrt(var_alpha ** 2 + (g_dc1 ** 2 * var_dc ** 2) + s1 ** 2)
    c[:, 18] = g_bd2 * mu_alpha + mu_dc
    c[:, 19] = np.sqrt((g_bd2 ** 2 * var_alpha ** 2) + var_dc ** 2 + s2 ** 2)
    c[:, 20], c[:, 21], c[:, 22] = ORDER_DC_BOUND_DRIFT, 2, 1
    return c


def canonical_drift_alpha(params):
    """[mu_drift, mu_alpha, beta, tau, eta, dc, var_alpha, gamma_bd1, gamma_dr2, sigma1, sigma2] -> (B, 24)."""
    p = np.atleast_2d(np.asarray(params, dtype=np.float64))
    mu_drift, mu_alpha, beta, tau, eta, dc, var_alpha, g_bd1, g_dr2, s1, s2 = p.T
    c = _canon(p.shape[0])
    c[:, 0], c[:, 1] = mu_drift, eta
    c[:, 2], c[:, 3] = mu_alpha, var_alpha
    c[:, 4] = dc
    c[:, 6], c[:, 7] = beta, tau
    c[:, 8], c[:, 9], c[:, 11] = 1.0, g_bd1, s1           # EEG1 = N(1*drift_t + gamma_bd1*bound_t, sigma1)
    c[:, 14], c[:, 15], c[:, 17] = g_dr2, 1.0, s2         # EEG2 = N(gamma_dr2*drift_t + 1*bound_t, sigma2)
    c[:, 20], c[:, 21], c[:, 22] = ORDER_DRIFT_BOUND_DC, 2, 1
    return c


def canonical_alpha_standardised(params):
    """single_trial_alpha.py: [drift, mu_alpha, beta, ter, var_alpha, dc, sigma1] -> (B, 24), one channel."""
    p = np.atleast_2d(np.asarray(params, dtype=np.float64))
    drift, mu_alpha, beta, ter, var_alpha, dc, s1 = p.T
    c = _canon(p.shape[0])
    c[:, 0] = drift
    c[:, 2], c[:, 3] = mu_alpha, var_alpha
    c[:, 4] = dc
    c[:, 6], c[:, 7] = beta, ter
    c[:, 9], c[:, 11] = 1.0, s1                            # temp = N(1*bound_t, sigma1)
    c[:, 12] = 1 * mu_alpha                                # single_trial_alpha.py:112-114
    c[:, 13] = np.sqrt(var_alpha ** 2 + s1 ** 2)
    c[:, 20], c[:, 21], c[:, 22] = ORDER_DRIFT_BOUND_DC, 1, 1
    return c


def _simulate(canon, n_trials, single, simulator, dt, max_steps, **kw):
    sim = simulator if simulator is not None else default_simulator()
    out = sim.simulate(_capi.MODEL_GENERAL, canon, int(n_trials), dt, int(max_steps), **kw)
    return out[0] if single else out


def simulate_trials_drift_dc5(params, n_trials, simulator=None, dt=.01, max_steps=400., **kw):
    """single_trial_drift_dc5.py:142-154 -> (n_trials, 3); (B, 11) parameters give (B, n_trials, 3)."""
    return _simulate(canonical_drift_dc5(params), n_trials, np.ndim(params) == 1, simulator, dt, max_steps, **kw)


def simulate_trials_drift_dc4(params, n_trials, simulator=None, dt=.01, max_steps=400., **kw):
    """single_trial_drift_dc4.py:134-146: the same channels, not standardised."""
    return _simulate(canonical_drift_dc5(params, standardise=False), n_trials, np.ndim(params) == 1, simulator, dt,
                     max_steps, **kw)


def simulate_trials_alpha_dc(params, n_trials, simulator=None, dt=.01, max_steps=400., **kw):
    """single_trial_alpha_dc.py:164-176 -> (n_trials, 3)."""
    return _simulate(canonical_alpha_dc(params), n_trials, np.ndim(params) == 1, simulator, dt, max_steps, **kw)


def simulate_trials_drift_alpha(params, n_trials, simulator=None, dt=.01, max_steps=400., **kw):
    """single_trial_drift_alpha.py:140-152 -> (n_trials, 3)."""
    return _simulate(canonical_drift_alpha(params), n_trials, np.ndim(params) == 1, simulator, dt, max_steps, **kw)


def simulate_trials_alpha_standardised(params, n_trials, simulator=None, dt=.01, max_steps=400., **kw):
    """single_trial_alpha.py:124-135 -> (n_trials, 2): (choicert, standardised extdata1)."""
    out = _simulate(canonical_alpha_standardised(params), n_trials, np.ndim(params) == 1, simulator, dt, max_steps, **kw)
    return np.ascontiguousarray(out[..., :2])
