"""Drop-in for the data-generation block of the reference's ``alpha_not_scaled.py`` (lines 52-131):
100 simulated participants x 100 trials of a DDM with trial-to-trial drift variability, plus one
external correlate of the boundary per participant, collected in the ``genparam`` dict the script
saves to ``data/alpha_not_scaled_test{test_num}.mat`` (and that ``basic_ddm_dc_pyjags.py`` /
``basic_ddm_dc_pystan2.py`` build the same way).

The reference draws the participant parameters from NumPy's legacy global state after
``np.random.seed(2021)`` (:62-88) -- reproduced here draw for draw, so the parameters are the
reference's own -- and then calls ``phju.simulratcliff`` per participant (:95-97), a pure-numpy
rejection sampler of the continuous-time process (~92 us/trial).  Here all participants are one
launch: by default of the same exact sampler on the GPU (``ddm_simulate_exact``, ``exact=True``), or
of the Euler-Maruyama kernel (DDM_MODEL_ETA: drift_trial ~ N(delta, deltatrialsd)) at a fine step,
dt = 1e-4 by default (``exact=False``; a different algorithm).  Streams differ from NumPy's,
so parity is distributional: tests compare against samples of the reference's sampler.  The JAGS fit
and the plots are out of scope.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .simulator import default_simulator

SIGMA_BY_TEST = {1: .5, 2: .1, 3: .01, 4: .2}  # alpha_not_scaled.py:72-80


def draw_participants(nparts=100, test_num=2, seed=2021):
    """alpha_not_scaled.py:62-88: per-participant generating parameters (legacy NumPy stream)."""
    rs = np.random.RandomState(seed)
    ndt = rs.uniform(.15, .6, size=nparts)
    alpha = rs.uniform(.8, 1.4, size=nparts)
    var_alpha = (1 / 12) * (1.4 - .8) ** 2
    beta = rs.uniform(.3, .7, size=nparts)
    delta = rs.uniform(-4, 4, size=nparts)
    varsigma = rs.uniform(.8, 1.4, size=nparts)
    deltatrialsd = rs.uniform(0, 2, size=nparts)
    sigma = SIGMA_BY_TEST[test_num]
    if nparts > 17:  # "Fix parameters across simulations" (:83-88)
        ndt[17], alpha[17], beta[17], delta[17], varsigma[17], deltatrialsd[17] = .4, 1.2, .5, 3.5, 1.2, 1
    return dict(ndt=ndt, alpha=alpha, beta=beta, delta=delta, varsigma=varsigma, deltatrialsd=deltatrialsd,
                sigma=sigma, var_alpha=var_alpha), rs


def generate_data(test_num=2, nparts=100, ntrials=100, seed=2021, simulator=None, dt=1e-4, max_time=20.,
                  sim_seed=None, exact=True):
    """alpha_not_scaled.py:52-131 -> the ``genparam`` dict (same keys, shapes and dtypes)."""
    g, rs = draw_participants(nparts, test_num, seed)
    sim = simulator if simulator is not None else default_simulator()
    if exact:
        # phju.simulratcliff(N=ntrials, Alpha, Tau, Nu, Beta, Eta, Varsigma) per participant (:95-97)
        zero = np.zeros(nparts)
        params = np.stack([g['alpha'], g['ndt'], g['delta'], g['beta'], zero, zero, g['deltatrialsd'], g['varsigma']], axis=-1)
        y = sim.simulate_exact(params, ntrials, seed=sim_seed).reshape(-1)
        rt, choice = np.abs(y), np.sign(y)
    else:
        # simulratcliff clips the mean drift to +-5 (pyhddmjagsutils.py:102-103); |delta| <= 4 here anyway
        nu = np.clip(g['delta'], -5, 5)
        params = np.stack([nu, g['alpha'], g['beta'], g['ndt'], g['deltatrialsd'], g['varsigma']], axis=-1)
        out = sim.simulate(_capi.MODEL_ETA, params, ntrials, dt, int(round(max_time / dt)), seed=sim_seed)
        rt = out[..., 0].reshape(-1)
        choice = out[..., 1].reshape(-1)
    N = ntrials * nparts
    # external data measured per participant (:103-106)
    if test_num != 4:
        extdata = rs.normal(loc=1 * g['alpha'], scale=g['sigma'])
    else:
        extdata = rs.normal(loc=1, scale=g['sigma'], size=nparts)
    genparam = dict(g)
    genparam['prop_cog_var'] = g['var_alpha'] / (g['var_alpha'] + g['sigma'] ** 2)
    genparam['rt'] = rt
    genparam['acc'] = (choice + 1) / 2
    genparam['y'] = choice * rt
    genparam['extdata'] = extdata
    genparam['participant'] = np.repeat(np.arange(1, nparts + 1, dtype=np.float64), ntrials)
    genparam['nparts'] = nparts
    genparam['ntrials'] = ntrials
    genparam['N'] = N
    return genparam
