"""CPU checks of the analytic oracles (check #2's targets): the Navarro-Fuss WFPT code against
closed forms and the reference's own exact sampler, and the exact law of the discrete Euler
chain against the reference loop."""
import os

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT
from oracle import euler_chain, wfpt

RATCLIFF = os.path.join(ROOT, "tests", "golden", "simulratcliff_samples.npz")


def test_wfpt_density_integrates_to_absorption_probability():
    for (v, a, w) in [(1.0, 1.5, 0.5), (-2.0, 1.0, 0.3), (0.0, 2.0, 0.6), (3.0, 0.8, 0.5)]:
        t = np.linspace(1e-6, 30.0, 400001)
        f = wfpt.pdf_lower(t, v, a, w)
        mass = np.trapezoid(f, t)
        assert abs(mass - wfpt.prob_lower(v, a, w)) < 2e-4
        # CDF is the integral of the density
        i = 40000
        assert abs(np.trapezoid(f[: i + 1], t[: i + 1]) - wfpt.cdf_lower(t[i], v, a, w)[0]) < 2e-4


def test_wfpt_small_and_large_time_series_agree():
    u = np.array([0.2, 0.3, 0.34, 0.36, 0.5])
    assert np.allclose(wfpt._f1_small(u, 0.37), wfpt._f1_large(u, 0.37), rtol=1e-8, atol=1e-12)


def test_wfpt_against_reference_exact_sampler():
    """simulratcliff (pyhddmjagsutils.py:47-176) samples the continuous-time law exactly; the
    golden samples were drawn by the unmodified reference (make_golden_ratcliff.py)."""
    z = np.load(RATCLIFF)
    for name in ("fast_unbiased", "slow_biased", "dc_scaled_pair_a", "dc_scaled_pair_b"):
        alpha, tau, nu, beta, eta, vs = z[f"{name}__params"]
        y = z[f"{name}__y"]
        s = np.sign(y) * (np.abs(y) - tau)     # signed decision time
        res = stats.kstest(s, lambda x: wfpt.signed_rt_cdf(x, nu, alpha, beta, vs))
        assert res.pvalue > 1e-3, (name, res)
        assert abs(np.mean(y > 0) - wfpt.ddm_prob_upper(nu, alpha, beta, vs)) < 4 * 0.5 / np.sqrt(y.size)


def test_dc_scaling_identity_in_reference_samples():
    """(boundary, drift, dc) and (2x, 2x, 2x) have the same law (Basic_DDM_simulations.py:164-209)."""
    z = np.load(RATCLIFF)
    a, b = z["dc_scaled_pair_a__y"], z["dc_scaled_pair_b__y"]
    assert stats.ks_2samp(a, b).pvalue > 1e-3


def test_euler_chain_pmf_is_a_distribution_and_matches_the_reference_loop(oracle):
    for params, dt, ms in [([3.0, 1.5, 0.5, 0.4, 1.0], 0.01, 400), ([-1.0, 0.9, 0.35, 0.2, 1.3], 0.01, 400),
                           ([0.05, 4.0, 0.5, 0.3, 0.3], 0.01, 400)]:
        pu, pl, pt = euler_chain.first_passage_pmf(params[0], params[1], params[2], params[4], dt, ms)
        assert abs(pu.sum() + pl.sum() + pt - 1) < 1e-9 and pu.min() >= 0 and pl.min() >= 0
        n = 30000
        t = oracle.simulate_mt(0, params, n, seed=321, dt=dt, max_steps=float(ms))
        signed = t.choice.astype(np.int64) * t.n_steps
        support, cdf = euler_chain.signed_step_cdf(pu, pl, pt)
        ecdf = np.searchsorted(np.sort(signed), support, side="right") / n
        d = np.max(np.abs(ecdf - cdf))
        assert d < 1.63 / np.sqrt(n), (params, d)         # KS critical value at alpha = 0.01
        assert abs((t.choice == 0).mean() - pt) < 4 * np.sqrt(max(pt, 1e-4) / n)


def test_euler_chain_converges_to_wfpt():
    drift, bound, beta, dc = 1.0, 1.2, 0.5, 1.0
    dmax = []
    for dt, ms in [(0.01, 400), (0.0025, 1600)]:
        pu, pl, pt = euler_chain.first_passage_pmf(drift, bound, beta, dc, dt, ms, grid=1200)
        tt = np.arange(1, ms + 1) * dt
        cu = np.cumsum(pu[1:])
        cw = wfpt.ddm_cdf(tt, +1, drift, bound, beta, dc)
        dmax.append(np.max(np.abs(cu - cw)))
    assert dmax[1] < dmax[0] and dmax[1] < 0.03
    # discretisation bias is first order in sqrt(dt): visible at dt = .01 (why KS targets the chain)
    assert dmax[0] > 0.01


def test_continuity_corrected_wfpt_matches_reference_loop(oracle):
    """The reference loop (dt = 1e-3) against WFPT with the discretely-monitored-barrier correction
    c = 0.5826*dc*sqrt(dt): this is the analytic target the GPU KS test uses."""
    p = [-2.0, 1.6, 0.4, 0.0, 1.4]
    n, dt = 20000, 1e-3
    t = oracle.simulate_mt(0, p, n, seed=5, dt=dt, max_steps=20000.0)
    s = t.sim_data[:, 1] * t.sim_data[:, 0]
    c = 0.5826 * p[4] * np.sqrt(dt)
    a2 = p[1] + 2 * c
    b2 = (p[1] * p[2] + c) / a2
    raw = stats.kstest(s, lambda x: wfpt.signed_rt_cdf(x, p[0], p[1], p[2], p[4])).statistic
    cor = stats.kstest(s, lambda x: wfpt.signed_rt_cdf(x, p[0], a2, b2, p[4])).statistic
    assert cor * np.sqrt(n) < 1.63 and raw > 2 * cor
