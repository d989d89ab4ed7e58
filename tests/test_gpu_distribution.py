"""Check #2: distribution-level parity of the CUDA simulator (its own Philox stream) against
(a) the exact law of the discrete Euler chain, (b) the analytic Navarro-Fuss Wiener
first-passage distribution at small dt, (c) the reference's exact sampler, (d) the reference
loop on its own stream (oracle), for every model variant; plus full-size property checks."""
import os

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT
from oracle import euler_chain, wfpt

pytestmark = pytest.mark.gpu
F_F32, F_STEPS = 2, 4
KS_01 = 1.63  # sqrt(n) * D critical value at alpha = 0.01


def _signed_steps(sim, model, params, n, **kw):
    out = sim.simulate(model, params, n, flags=F_STEPS, **kw)[0]
    steps = sim.last_steps(n).astype(np.int64)
    choice = out[:, 1].astype(np.int64) if model == 0 else np.sign(out[:, 0]).astype(np.int64)
    return choice * steps, out


@pytest.mark.parametrize("params", [[3.0, 1.5, 0.5, 0.4, 1.0], [-1.0, 0.9, 0.35, 0.2, 1.3], [0.3, 1.1, 0.6, 0.0, 0.6],
                                    [0.05, 4.0, 0.5, 0.3, 0.3]])
def test_ks_against_exact_discrete_chain(sim, params):
    """dt = .01 (the reference default): the target is the chain's own first-passage law."""
    n = 200_000
    signed, out = _signed_steps(sim, 0, params, n, seed=17, dataset_offset=0, dt=0.01, max_steps=400)
    pu, pl, pt = euler_chain.first_passage_pmf(params[0], params[1], params[2], params[4], 0.01, 400, grid=1500)
    support, cdf = euler_chain.signed_step_cdf(pu, pl, pt)
    ecdf = np.searchsorted(np.sort(signed), support, side="right") / n
    assert np.max(np.abs(ecdf - cdf)) * np.sqrt(n) < KS_01
    assert abs((out[:, 1] == 0).mean() - pt) < 4 * np.sqrt(max(pt, 1e-5) / n) + 1e-5
    # RTs live on the lattice n*dt + tau, computed in fp64 exactly as the reference does
    assert np.array_equal(out[:, 0], np.abs(signed) * 0.01 + params[3]) or np.array_equal(
        out[out[:, 1] != 0, 0], (np.abs(signed) * 0.01 + params[3])[out[:, 1] != 0])


def _corrected(params, dt):
    """Continuity correction of the discretely monitored barrier (Siegmund 1985; Broadie, Glasserman &
    Kou 1997): the Euler chain absorbed at (0, a) behaves like the continuous process absorbed at
    (-c, a + c), c = 0.5826 * dc * sqrt(dt), up to O(dt)."""
    c = 0.5826 * params[4] * np.sqrt(dt)
    a2 = params[1] + 2 * c
    return a2, (params[1] * params[2] + c) / a2


@pytest.mark.parametrize("params", [[1.0, 1.2, 0.5, 0.0, 1.0], [-2.0, 1.6, 0.4, 0.0, 1.4], [0.0, 0.8, 0.7, 0.0, 0.5]])
def test_ks_against_analytic_wiener_fpt(sim, params):
    """Constant-parameter DDM at dt = 1e-4: one-sample KS of the signed decision time against the
    Navarro-Fuss CDF with the reference's dc scaling (a = alpha/dc, v = drift/dc).  The Euler scheme's
    first-order (sqrt(dt)) barrier bias is removed with the standard continuity correction; the
    uncorrected distance must still be small and shrink with dt."""
    n, dt = 20_000, 1e-4
    out = sim.simulate(0, params, n, seed=23, dataset_offset=0, dt=dt, max_steps=200_000)[0]
    assert (out[:, 1] == 0).sum() == 0
    s = out[:, 1] * out[:, 0]
    a2, b2 = _corrected(params, dt)
    res = stats.kstest(s, lambda x: wfpt.signed_rt_cdf(x, params[0], a2, b2, params[4]))
    assert res.statistic * np.sqrt(n) < KS_01, res
    raw = stats.kstest(s, lambda x: wfpt.signed_rt_cdf(x, params[0], params[1], params[2], params[4]))
    assert raw.statistic < 0.025
    assert abs((out[:, 1] > 0).mean() - wfpt.ddm_prob_upper(params[0], a2, b2, params[4])) < 0.012
    # ten times coarser step: the uncorrected bias grows like sqrt(dt), the corrected fit still holds
    out2 = sim.simulate(0, params, n, seed=24, dataset_offset=0, dt=1e-3, max_steps=20_000)[0]
    s2 = out2[:, 1] * out2[:, 0]
    a3, b3 = _corrected(params, 1e-3)
    assert stats.kstest(s2, lambda x: wfpt.signed_rt_cdf(x, params[0], a3, b3, params[4])).statistic * np.sqrt(n) < KS_01 * 1.3


def test_against_reference_exact_sampler(sim):
    z = np.load(os.path.join(ROOT, "tests", "golden", "simulratcliff_samples.npz"))
    for name in ("fast_unbiased", "slow_biased"):
        alpha, tau, nu, beta, eta, vs = z[f"{name}__params"]
        y = z[f"{name}__y"]
        out = sim.simulate(0, [nu, alpha, beta, tau, vs], 20_000, seed=29, dataset_offset=0, dt=2e-4, max_steps=100_000)[0]
        mine = out[:, 0] * out[:, 1]
        assert stats.ks_2samp(mine, y).pvalue > 1e-3, name


MODEL_CASES = [
    (1, [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1], dict(dt=0.01, max_steps=400)),
    (1, [0.5, 0.2, 0.4, 0.3, 2.5, 1.2, 3.0], dict(dt=0.01, max_steps=400)),
    (1, [1.0, 1.2, 0.5, 0.3, 0.3, 1.0, 0.5], dict(dt=0.001, max_steps=4000)),   # simulate_trials_fine
    (2, [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1], dict(dt=0.01, max_steps=400)),
    (2, [-1.0, 1.0, 0.55, 0.2, 2.0, 0.3, 2.0], dict(dt=0.01, max_steps=400)),
    (3, [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1, 1.37], dict(dt=0.01, max_steps=400)),
    (4, [2.0, 1.0, 0.45, 0.4, 0.5, 1.1, 0.7], dict(dt=0.01, max_steps=400)),
    (0, [0.4, 1.3, 0.55, 0.25, 0.9], dict(dt=0.001, max_steps=4000)),
    (6, [3.5, 1.2, 0.5, 0.4, 1.0, 1.2], dict(dt=0.01, max_steps=400)),
    (6, [0.0, 1.0, 0.5, 0.3, 2.5, 1.0], dict(dt=0.01, max_steps=400)),
]


@pytest.mark.parametrize("model,params,kw", MODEL_CASES)
def test_two_sample_against_reference_loop(sim, oracle, model, params, kw):
    """GPU (Philox, fp32) vs the reference loop on its own MT19937 stream: both output columns."""
    n = 60_000
    g = sim.simulate(model, params, n, seed=31, dataset_offset=2, **kw)[0]
    r = oracle.simulate_mt(model, params, n, seed=77, **kw).sim_data
    if model in (0, 6):
        gs, rs = g[:, 0] * g[:, 1], r[:, 0] * r[:, 1]
    else:
        gs, rs = g[:, 0], r[:, 0]
        assert stats.ks_2samp(g[:, 1], r[:, 1]).pvalue > 1e-3            # ext-data column
        # joint structure: ext-data tracks the trial's latent, so it correlates with |choicert|
        cg, cr = np.corrcoef(np.abs(gs), g[:, 1])[0, 1], np.corrcoef(np.abs(rs), r[:, 1])[0, 1]
        assert abs(cg - cr) < 0.02
    assert stats.ks_2samp(gs, rs).pvalue > 1e-3
    assert abs((gs == 0).mean() - (rs == 0).mean()) < 0.005


def test_trialwise_distribution(sim, oracle):
    p = [3.2, 0.48, 0.41, 1.05]
    bounds = np.full(40_000, 1.4)
    g = sim.simulate_trialwise(np.zeros(bounds.size, np.int32), bounds, [p], seed=5)[:, 0]
    r = oracle.simulate_mt(5, p, bounds.size, 5, bound_in=bounds).sim_data[:, 0]
    assert stats.ks_2samp(g, r).pvalue > 1e-3


def test_full_size_config3_properties(sim):
    """BASELINE config 3: single_trial_alpha_not_scaled, 1024 datasets x 1000 trials."""
    from bayesflow_nddms_b200 import priors

    params = priors.draw_prior_batch("alpha", 1024, np.random.default_rng(2023))
    out = sim.simulate(1, params, 1000, seed=3, dataset_offset=0, flags=F_STEPS)
    steps = sim.last_steps(1024 * 1000).reshape(1024, 1000)
    st = sim.last_stats()
    assert out.shape == (1024, 1000, 2) and np.all(np.isfinite(out))
    assert st["total_steps"] == int(steps.sum()) and st["used_persistent"] == 1 and st["reject_cap_hits"] == 0
    ter = params[:, 3][:, None]
    rt = np.abs(out[..., 0])
    miss = out[..., 0] == 0
    assert np.array_equal(miss, (steps == 400) & miss) and st["n_timeouts"] == int(miss.sum())
    # |choicert| = ter + n*dt, in the reference's fp64 arithmetic
    assert np.array_equal(rt[~miss], (ter + steps * 0.01)[~miss])
    # ext-data = N(bound_trial, sigma1): per-dataset mean near the truncated-normal mean of the boundary
    mu, sd, sig = params[:, 1], params[:, 4], params[:, 6]
    tn_mean = np.array([stats.truncnorm.mean(-m / s, np.inf, loc=m, scale=s) for m, s in zip(mu, sd)])
    tn_var = np.array([stats.truncnorm.var(-m / s, np.inf, loc=m, scale=s) for m, s in zip(mu, sd)])
    zscore = (out[..., 1].mean(1) - tn_mean) / np.sqrt((tn_var + sig ** 2) / 1000)
    assert abs(zscore.mean()) < 0.2 and 0.8 < zscore.std() < 1.2


def test_full_size_sweep_properties(sim):
    """A 2e7-trial slice of the throughput sweep (BASELINE config 5) -- every trial written once,
    counters consistent, step-count profile as surveyed (mean ~258, ~0.4% timeouts)."""
    from bayesflow_nddms_b200 import priors

    B, N = 20_000, 1000
    params = priors.draw_prior_batch("sweep", B, np.random.default_rng(1))
    out = sim.simulate(0, params, N, seed=9, dataset_offset=0, dt=1e-3, max_steps=4000, flags=F_STEPS | F_F32)
    steps = sim.last_steps(B * N).reshape(B, N)
    st = sim.last_stats()
    assert st["total_steps"] == int(steps.sum(dtype=np.int64))
    assert st["n_timeouts"] == int((out[..., 1] == 0).sum()) and st["n_upper"] == int((out[..., 1] > 0).sum())
    assert np.array_equal(out[..., 0], (steps * 1e-3).astype(np.float32))
    assert np.all(steps[out[..., 1] == 0] == 4000) and steps.max() <= 4000
    mean = steps.mean()
    assert 230 < mean < 290 and 0.002 < (out[..., 1] == 0).mean() < 0.008
    # upper-boundary probability per dataset against the closed form (continuous limit)
    sel = np.where((out[:400, :, 1] == 0).mean(1) < 0.01)[0]          # datasets that (almost) never time out
    pu = np.array([wfpt.ddm_prob_upper(params[i, 0], *_corrected(params[i], 1e-3), params[i, 4]) for i in sel])
    emp = (out[sel, :, 1] > 0).mean(1)
    assert sel.size > 300 and np.mean(np.abs(pu - emp)) < 0.015 and np.max(np.abs(pu - emp)) < 0.08


def _np_hist(out, basic, n_bins, rt_max):
    if basic:
        rt, ch = out[..., 0].ravel().astype(np.float64), np.sign(out[..., 1].ravel())
    else:
        rt, ch = np.abs(out[..., 0].ravel()).astype(np.float64), np.sign(out[..., 0].ravel())
    b = np.floor(rt * (n_bins / rt_max)).astype(np.int64)
    ok = ch != 0
    over = ok & (b >= n_bins)
    up = np.bincount(b[ok & ~over & (ch > 0)], minlength=n_bins)
    lo = np.bincount(b[ok & ~over & (ch < 0)], minlength=n_bins)
    return up, lo, int((~ok).sum()), int(over.sum())


@pytest.mark.parametrize("model,prior,basic", [(0, "basic", True), (1, "alpha", False), (6, "eta", True)])
def test_device_histogram_matches_numpy_and_adds_over_shards(sim, model, prior, basic):
    """The on-device RT histogram (SURVEY 8d, C5's reduction) equals numpy's on the downloaded rows, for both
    row layouts and dtypes, and the histograms of two dataset shards add up to the whole batch's."""
    from bayesflow_nddms_b200 import priors

    params = priors.draw_prior_batch(prior, 300, np.random.default_rng(model + 20))
    kw = dict(dt=0.01, max_steps=150, seed=3)
    for flags in (0, F_F32):
        out = sim.simulate(model, params, 211, dataset_offset=10, flags=flags, **kw)
        h = sim.last_histogram(64, 1.2)
        up, lo, missing, over = _np_hist(out, basic, 64, 1.2)
        assert np.array_equal(h["upper"], up) and np.array_equal(h["lower"], lo)
        assert h["missing"] == missing == sim.last_stats()["n_timeouts"] and h["overflow"] == over and over > 0
        assert int(h["upper"].sum() + h["lower"].sum()) + missing + over == 300 * 211
    sim.simulate(model, params[:123], 211, dataset_offset=10, flags=F_F32, **kw)
    h1 = sim.last_histogram(64, 1.2)
    sim.simulate(model, params[123:], 211, dataset_offset=133, flags=F_F32, **kw)
    h2 = sim.last_histogram(64, 1.2)
    assert np.array_equal(h1["upper"] + h2["upper"], h["upper"]) and np.array_equal(h1["lower"] + h2["lower"], h["lower"])
    assert h1["missing"] + h2["missing"] == h["missing"]


def test_full_size_1e9_trials_histogram_checksum(sim):
    """BASELINE config 5 at full size on one GPU -- 1e6 datasets x 1000 trials, float32 rows resident in HBM (8 GB):
    nothing is downloaded; the device histogram must account for every trial, agree with the kernel's own
    counters, and equal the sum of the histograms of four dataset shards run separately (a checksum of
    checksums: Philox counters carry the global dataset index, so shards reproduce the batch)."""
    import torch

    from bayesflow_nddms_b200 import priors

    if torch.cuda.mem_get_info()[0] < 30e9:
        pytest.skip("needs 30 GB of free HBM")
    B, N = 1_000_000, 1000
    params = priors.draw_prior_batch("sweep", B, np.random.default_rng(5))
    kw = dict(dt=1e-3, max_steps=4000, seed=17)
    sim.run(0, params, N, dataset_offset=0, flags=F_F32, **kw)
    st = sim.last_stats()
    h = sim.last_histogram(401, 4.01)   # a response at the 4000th step has rt = 4.0
    assert st["n_trials"] == B * N
    assert int(h["upper"].sum()) == st["n_upper"] and h["missing"] == st["n_timeouts"] and h["overflow"] == 0
    assert int(h["upper"].sum() + h["lower"].sum()) + h["missing"] == B * N
    assert 230 * B * N < st["total_steps"] < 290 * B * N
    # mean RT from the histogram (bin centres) against total_steps * dt (tau = 0 in the sweep), timeouts excluded
    centres = 0.5 * (h["edges"][:-1] + h["edges"][1:])
    responded = B * N - h["missing"]
    mean_rt = float(((h["upper"] + h["lower"]) * centres).sum()) / responded
    assert abs(mean_rt - (st["total_steps"] - 4000 * h["missing"]) * 1e-3 / responded) < 0.006
    up = np.zeros_like(h["upper"]); lo = np.zeros_like(h["lower"]); steps = 0
    for k in range(4):
        a, b = k * B // 4, (k + 1) * B // 4
        sim.run(0, params[a:b], N, dataset_offset=a, flags=F_F32, **kw)
        hk = sim.last_histogram(401, 4.01)   # a response at the 4000th step has rt = 4.0
        up += hk["upper"]; lo += hk["lower"]; steps += sim.last_stats()["total_steps"]
    assert np.array_equal(up, h["upper"]) and np.array_equal(lo, h["lower"]) and steps == st["total_steps"]


def test_alpha_not_scaled_data_generation(sim):
    """alpha_not_scaled.py:52-131 with the GPU generator in place of simulratcliff: the genparam dict,
    and participant 17 (fixed parameters, drift variability eta = 1) against the reference sampler's
    own samples (different algorithm, distribution-level parity)."""
    from bayesflow_nddms_b200 import alpha_not_scaled as m

    acc = {}
    for exact in (True, False):      # the GPU exact sampler (default) and the fine-step Euler kernel
        g = m.generate_data(test_num=2, simulator=sim, sim_seed=3, exact=exact)
        assert g['N'] == 10000 and g['rt'].shape == g['acc'].shape == g['y'].shape == g['participant'].shape == (10000,)
        assert g['extdata'].shape == (100,) and set(np.unique(g['acc'])) <= {0.0, 0.5, 1.0}
        assert np.array_equal(g['y'], (2 * g['acc'] - 1) * g['rt']) and np.all(g['rt'] >= np.repeat(g['ndt'], 100))
        assert abs(np.corrcoef(g['extdata'], g['alpha'])[0, 1]) > 0.7           # sigma = .1 vs sd(alpha) = .17
        assert abs(g['prop_cog_var'] - 0.03 / 0.04) < 1e-12
        acc[exact] = g['acc'].reshape(100, 100).mean(1)
    assert np.corrcoef(acc[True], acc[False])[0, 1] > 0.9                      # same participants, same accuracies
    z = np.load(os.path.join(ROOT, "tests", "golden", "simulratcliff_samples.npz"))
    y = z["participant17_eta__y"]
    alpha, tau, nu, beta, eta, vs = z["participant17_eta__params"]
    out = sim.simulate(6, [nu, alpha, beta, tau, eta, vs], 40_000, dt=1e-4, max_steps=200_000, seed=4, dataset_offset=0)[0]
    assert stats.ks_2samp(out[:, 0] * out[:, 1], y).pvalue > 1e-3
    assert abs((out[:, 1] > 0).mean() - (y > 0).mean()) < 0.015


@pytest.mark.parametrize("params", [[1.0, 1.2, 0.5, 0.3, 1.0], [-0.6, 1.6, 0.35, 0.2, 0.8]])
def test_ks_against_exact_discrete_chain_at_2e7_trials(sim, params):
    """Distribution-exactness where it can be tested hardest: 2e7 trials of one constant-parameter DDM
    against the exact first-passage law of the Euler chain (quadrature error ~1e-6).  The KS critical
    value at this n is 3.6e-4 in CDF distance; the chi-square over the step bins must be unremarkable too."""
    n = 20_000_000
    out = sim.simulate(0, params, n, seed=41, dataset_offset=0, dt=0.01, max_steps=400, flags=F_STEPS | F_F32)[0]
    steps = sim.last_steps(n).astype(np.int64)
    signed = out[:, 1].astype(np.int64) * steps
    counts = np.bincount(signed + 400, minlength=801).astype(np.float64)
    pu, pl, pt = euler_chain.first_passage_pmf(params[0], params[1], params[2], params[4], 0.01, 400, grid=2000)
    support, cdf = euler_chain.signed_step_cdf(pu, pl, pt)
    ecdf = np.cumsum(counts) / n
    assert np.max(np.abs(ecdf - cdf)) * np.sqrt(n) < KS_01
    pmf = np.diff(np.concatenate([[0.0], cdf]))
    big = pmf * n >= 50
    chi2 = np.sum((counts[big] - n * pmf[big]) ** 2 / (n * pmf[big]))
    dof = int(big.sum()) - 1
    assert stats.chi2.sf(chi2, dof) > 1e-4, (chi2, dof)
