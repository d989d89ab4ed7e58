"""Device-side batched prior: the kernel against a Python restatement of its own Philox
mapping, against the reference's distributions, and wired into the generative-model dict."""
import numpy as np
import pytest
from scipy import stats
from scipy.special import ndtr, ndtri

pytestmark = pytest.mark.gpu


def _uniforms(oracle, seed, draw, block):
    w = oracle.philox4x32_10((2, block, draw & 0xFFFFFFFF, draw >> 32), (seed & 0xFFFFFFFF, seed >> 32))
    ua = ((int(w[0]) >> 5) * 67108864.0 + (int(w[1]) >> 6) + 0.5) / 9007199254740992.0
    ub = ((int(w[2]) >> 5) * 67108864.0 + (int(w[3]) >> 6) + 0.5) / 9007199254740992.0
    return ua, ub


def _tn(u, mean, sd, low, upp):
    fa, fb = ndtr((low - mean) / sd), ndtr((upp - mean) / sd)
    return min(max(mean + sd * ndtri(fa + u * (fb - fa)), low), upp)


def test_prior_kernel_matches_its_restatement(sim, oracle):
    seed, off = 0xABCDEF0123, 5_000_000_000          # draw indices beyond 2^32 use the high counter word
    got = sim.draw_prior("alpha_scale", 48, seed=seed, draw_offset=off)
    assert got.shape == (48, 8)
    for i in range(48):
        d = off + i
        ua, ub = _uniforms(oracle, seed, d, 0)
        want = [0.0 + 2.0 * np.sqrt(-2 * np.log(ua)) * np.cos(2 * np.pi * ub)]
        want.append(_tn(_uniforms(oracle, seed, d, 1)[0], 1.0, 0.5, 0.0, 10.0))
        u0, u1 = _uniforms(oracle, seed, d, 2)
        u2 = _uniforms(oracle, seed, d, 8)[0]
        want.append(sorted([u0, u1, u2])[1])
        want.append(_tn(_uniforms(oracle, seed, d, 3)[0], 0.5, 0.25, 0.0, 1.5))
        want.append(_tn(_uniforms(oracle, seed, d, 4)[0], 1.0, 0.5, 0.0, 3.0))
        want.append(_tn(_uniforms(oracle, seed, d, 5)[0], 1.0, 0.5, 0.0, 10.0))
        want.append(5.0 * _uniforms(oracle, seed, d, 6)[0])
        want.append(2.0 * _uniforms(oracle, seed, d, 7)[0])
        assert np.allclose(got[i], want, rtol=1e-11, atol=1e-12), i
    # column subsets / orders of the other families
    b = sim.draw_prior("basic", 48, seed=seed, draw_offset=off)
    assert np.array_equal(b[:, :4], got[:, :4]) and np.allclose(b[:, 4], [
        _tn(_uniforms(oracle, seed, off + i, 4)[0], 1.0, 0.5, 0.0, 10.0) for i in range(48)], rtol=1e-11)
    sw = sim.draw_prior("sweep", 48, seed=seed, draw_offset=off)
    assert np.all(sw[:, 3] == 0) and np.array_equal(sw[:, [0, 1, 2, 4]], b[:, [0, 1, 2, 4]])
    # shardable: the draw index keys the stream
    lo = sim.draw_prior("basic", 20, seed=seed, draw_offset=off)
    hi = sim.draw_prior("basic", 28, seed=seed, draw_offset=off + 20)
    assert np.array_equal(np.concatenate([lo, hi]), b)


def test_prior_marginals_match_reference_distributions(sim):
    tn = lambda m, s, lo, hi: stats.truncnorm((lo - m) / s, (hi - m) / s, loc=m, scale=s)  # noqa: E731
    n = 200_000
    targets = {
        "basic": [stats.norm(0, 2), tn(1, .5, 0, 10), stats.beta(2, 2), tn(.5, .25, 0, 1.5), tn(1, .5, 0, 10)],
        "alpha": [stats.norm(0, 2), tn(1, .5, 0, 10), stats.beta(2, 2), tn(.5, .25, 0, 1.5), tn(1, .5, 0, 3), tn(1, .5, 0, 10),
                  stats.uniform(0, 5)],
        "alpha_scale": [None] * 7 + [stats.uniform(0, 2)],
        "eta": [None] * 4 + [tn(1, .5, 0, 3), tn(1, .5, 0, 10)],
        "evidence": [None] * 4 + [tn(1, .5, 0, 10), stats.uniform(0, 5)],
    }
    for name, dists in targets.items():
        p = sim.draw_prior(name, n, seed=11, draw_offset=0)
        assert p.shape == (n, len(dists)) and np.all(np.isfinite(p))
        for j, d in enumerate(dists):
            if d is not None:
                assert stats.kstest(p[:, j], d.cdf).statistic * np.sqrt(n) < 1.63, (name, j)
    p = sim.draw_prior("alpha", n, seed=11, draw_offset=0)
    c = np.corrcoef(p.T)
    assert np.max(np.abs(c - np.eye(7))) < 0.01          # independent columns


def test_generative_model_with_device_prior(sim):
    import torch

    from bayesflow_nddms_b200 import basic_ddm_dc as m0
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

    sim.dataset_counter = 4242
    d = m0.generative_model(32, sim, device_prior=True)
    n = d['sim_non_batchable_context']
    assert d['prior_draws'].shape == (32, 5) and d['sim_data'].shape == (32, n, 2) and sim.dataset_counter == 4242 + 32
    # the simulation used exactly these draws, keyed by the same global dataset indices
    again = sim.simulate(0, d['prior_draws'], n, dataset_offset=4242)
    assert np.array_equal(again, d['sim_data'])
    c = m0.configurator(d)
    assert c['parameters'].shape == (32, 5) and c['summary_conditions'].shape == (32, n, 2)
    dd = m1.generative_model(16, sim, device=True, device_prior=True)
    t = torch.from_dlpack(dd['sim_data'])
    assert t.is_cuda and tuple(t.shape) == (16, dd['sim_non_batchable_context'], 2) and dd['prior_draws'].shape == (16, 7)
    assert torch.isfinite(t).all()


@pytest.mark.gpu
@pytest.mark.parametrize("prior,cols", [("basic", 5), ("alpha", 7), ("eta", 6)])
def test_training_batch_is_the_three_calls(sim, prior, cols):
    """ddm_training_batch (one FFI crossing, one stream synchronisation) returns the draws and the batch of
    ddm_draw_prior + ddm_run + ddm_last_output_dlpack, bit for bit, and advances the dataset counter like them."""
    import torch

    from bayesflow_nddms_b200 import _capi

    sim.dataset_counter = 9000
    draws, batch = sim.training_batch(prior, 48, 333, 0.01, 400)
    t = torch.from_dlpack(batch).cpu().numpy()
    assert draws.shape == (48, cols) and t.shape == (48, 333, 2) and t.dtype == np.float32 and sim.dataset_counter == 9048
    sim.dataset_counter = 9000
    draws3 = sim.draw_prior(prior, 48)
    sim.run_uploaded(333, 0.01, 400, flags=_capi.FLAG_OUT_F32)
    t3 = torch.from_dlpack(sim.last_output_dlpack()).cpu().numpy()
    assert np.array_equal(draws, draws3) and np.array_equal(t, t3)
    # the batch is the simulation of exactly these draws
    again = sim.simulate(_capi.PRIORS[prior][0] if prior != "basic" else 0, draws, 333, dataset_offset=9000, flags=_capi.FLAG_OUT_F32)
    assert np.array_equal(again, t)
    # no host copy of the draws
    sim.dataset_counter = 9000
    none, batch = sim.training_batch(prior, 48, 333, 0.01, 400, to_host=False)
    assert none is None and np.array_equal(torch.from_dlpack(batch).cpu().numpy(), t)
    with pytest.raises(ValueError):
        sim.training_batch("evidence", 4, 10)
