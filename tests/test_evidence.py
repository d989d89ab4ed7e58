"""Evidence-path variants (retired model zoo): oracle pinned to the reference (CPU), CUDA
kernels against the oracle (GPU)."""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "reference_evidence.npz")
F_F32 = 2


def _cases():
    z = np.load(GOLD)
    for name in sorted({k.split("__")[0] for k in z.files}):
        n, seed, mode, n_obs = (int(v) for v in z[f"{name}__meta"])
        p = z[f"{name}__params"]
        if p.size == 5:
            p = np.append(p, 0.001)
        yield name, p, n, seed, mode, n_obs, z[f"{name}__out"]


def test_oracle_reproduces_reference_evidence_outputs_bit_exact(oracle):
    seen = set()
    for name, p, n, seed, mode, n_obs, ref in _cases():
        out, ns, _ = oracle.simulate_evidence(p, n, n_obs, mode, mt_seed=seed)
        assert np.array_equal(out.view(np.uint64), ref.view(np.uint64)), name
        seen.add((mode, n_obs))
        # structure: the path is held (up to noise) after the crossing step; z-scored rows have mean 0, sd 1
        if mode == 1:
            assert np.allclose(out[:, 2:].mean(1), 0, atol=1e-12) and np.allclose(out[:, 2:].std(1), 1, atol=1e-12)
    assert seen == {(1, 200), (2, 200), (1, 400)}


def test_oracle_buffer_source_matches_mt(oracle):
    for name, p, n, seed, mode, n_obs, ref in _cases():
        a, ns, _ = oracle.simulate_evidence(p, n, n_obs, mode, mt_seed=seed)
        normals = oracle.mt_normals(seed, int(ns.sum()) + n * n_obs + 8)
        b, ns2, cons = oracle.simulate_evidence(p, n, n_obs, mode, normals=normals)
        assert np.array_equal(a, b) and np.array_equal(cons, ns + n_obs), name


@pytest.mark.gpu
def test_gpu_fp64_shared_increments_reproduce_reference(sim, oracle):
    for name, p, n, seed, mode, n_obs, ref in _cases():
        _, ns, _ = oracle.simulate_evidence(p, n, n_obs, mode, mt_seed=seed)
        normals = oracle.mt_normals(seed, int(ns.sum()) + n * n_obs + 8)
        off = np.zeros(n, np.int64)
        off[1:] = np.cumsum(ns + n_obs)[:-1]
        sim.set_normals_debug(normals, off)
        try:
            out = sim.simulate_evidence(p, n, n_obs, mode, precision=64, seed=1, dataset_offset=0)[0]
            st = sim.last_stats()
            f32 = sim.simulate_evidence(p, n, n_obs, mode, precision=64, seed=1, dataset_offset=0, flags=F_F32)[0]
        finally:
            sim.set_normals_debug(None, None)
        assert np.array_equal(out.view(np.uint64), ref.view(np.uint64)), name
        assert st["total_steps"] == int(ns.sum()) and st["debug_overruns"] == 0
        assert np.array_equal(f32, ref.astype(np.float32)), name


@pytest.mark.gpu
@pytest.mark.parametrize("mode,n_obs,params", [
    (1, 200, [3.0, 1.0, 0.5, 0.4, 1.0, 1.0]), (1, 200, [0.2, 2.5, 0.45, 0.3, 0.6, 0.3]), (2, 200, [3.0, 1.0, 0.5, 0.4, 1.0, 1.0]),
    (2, 200, [-0.3, 2.0, 0.55, 0.35, 0.7, 0.5]), (1, 400, [3.0, 1.0, 0.5, 0.4, 1.0, 0.001]), (0, 37, [1.0, 1.5, 0.5, 0.2, 1.0, 0.2]),
    (1, 768, [0.5, 2.2, 0.5, 0.3, 0.8, 0.5])])
def test_gpu_fp32_production_vs_fp64_validation_and_oracle(sim, oracle, mode, n_obs, params):
    """Same Philox stream: the production warp kernel against the fp64 validation kernel and the
    oracle on the ideal stream.  Steps/choices equal except boundary ties; paths within fp32 tolerance."""
    n = 333                                   # not a multiple of 32: ragged last tile
    a = sim.simulate_evidence(params, n, n_obs, mode, seed=5, dataset_offset=9)[0]
    st = sim.last_stats()
    b = sim.simulate_evidence(params, n, n_obs, mode, seed=5, dataset_offset=9, precision=64)[0]
    o, ns, _ = oracle.simulate_evidence(params, n, n_obs, mode, philox_seed=5, dataset=9)
    assert a.shape == (n, 2 + n_obs) and np.all(np.isfinite(a))
    same64 = (b[:, 0] == o[:, 0]) & (b[:, 1] == o[:, 1])
    assert same64.mean() > 0.995
    assert np.allclose(b[same64, 2:], o[same64, 2:], rtol=0, atol=1e-9 if mode != 2 else 1e-6)
    same = (a[:, 0] == b[:, 0]) & (a[:, 1] == b[:, 1])
    assert same.mean() >= 0.98 and st["used_persistent"] == 1
    if mode != 2:  # dataset-level statistics couple every trial to the tie trials
        tol = 2e-3 if mode == 1 else 2e-5 * (1 + abs(params[1]))
        assert np.max(np.abs(a[same, 2:] - b[same, 2:])) < tol
    else:
        assert np.median(np.abs(a[:, 2:] - b[:, 2:])) < 5e-3
    if mode == 1:
        assert np.allclose(a[:, 2:].mean(1), 0, atol=1e-4) and np.allclose(a[:, 2:].std(1), 1, atol=1e-3)


@pytest.mark.gpu
def test_gpu_evidence_module_signatures_and_batching(sim):
    import torch

    from bayesflow_nddms_b200 import basic_ddm_dc_evidence as m

    p = m.draw_prior()
    assert p.shape == (6,)
    out = m.simulate_trials(p, 50, sim)
    assert out.shape == (50, 202) and out.dtype == np.float64
    assert m.simulate_trials_evidence2(p, 50, sim).shape == (50, 202)
    assert m.simulate_trials_no_noise2(m.draw_prior_no_noise(), 50, sim).shape == (50, 402)
    rt, choice, path = m.diffusion_trial(simulator=sim)
    assert rt >= 0.4 and choice in (-1, 0, 1) and path.shape == (200,)
    P = m.batch_draw_prior(16)
    host = m.simulate_trials(P, 64, sim, seed=3, dataset_offset=0, flags=F_F32)
    dev = m.simulate_trials(P, 64, sim, seed=3, dataset_offset=0, flags=F_F32, device=True)
    t = torch.from_dlpack(dev)
    assert tuple(t.shape) == (16, 64, 202) and np.array_equal(t.cpu().numpy(), host)
    # sharding invariance: datasets keyed by their global index
    lo = m.simulate_trials(P[:5], 64, sim, seed=3, dataset_offset=0, flags=F_F32)
    hi = m.simulate_trials(P[5:], 64, sim, seed=3, dataset_offset=5, flags=F_F32)
    assert np.array_equal(np.concatenate([lo, hi]), host)
    with pytest.raises(ValueError):
        sim.simulate_evidence(P, 10, 769)
