"""The general two-latent / two-channel model (DDM_MODEL_GENERAL) behind the retired zoo's
single_trial_drift_dc5 / _dc4 / single_trial_alpha_dc simulators: oracle pinned to the reference (CPU),
CUDA kernels against the oracle (GPU)."""
import os

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "reference_two_channel.npz")
F_F32, F_STEPS, F_GENERIC = 2, 4, 8


def _canon(variant, p):
    from bayesflow_nddms_b200 import two_channel as tc

    return {0: tc.canonical_drift_dc5, 1: lambda q: tc.canonical_drift_dc5(q, standardise=False), 2: tc.canonical_alpha_dc,
            3: tc.canonical_drift_alpha, 4: tc.canonical_alpha_standardised}[variant](p)[0]


def _cases():
    z = np.load(GOLD)
    for name in sorted({k.split("__")[0] for k in z.files}):
        n, seed, variant = (int(v) for v in z[f"{name}__meta"])
        yield name, _canon(variant, z[f"{name}__params"]), n, seed, z[f"{name}__out"]


def test_oracle_reproduces_reference_two_channel_outputs_bit_exact(oracle):
    n_cases = 0
    for name, canon, n, seed, ref in _cases():
        t = oracle.simulate_mt(7, canon, n, seed)
        got = np.ascontiguousarray(t.sim_data[:, :ref.shape[1]])       # the one-channel script stacks two columns
        assert got.shape == ref.shape, name
        assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), name
        n_cases += 1
    assert n_cases == 8


def test_canonical_mapping_layout():
    from bayesflow_nddms_b200 import two_channel as tc

    c = tc.canonical_drift_dc5([1.5, 1.2, 0.5, 0.4, 1.0, 1.0, 0.5, 0.7, -0.4, 0.3, 0.6])
    assert c.shape == (1, 24) and c[0, 3] == 0 and c[0, 1] == 1.0 and c[0, 21] == 2 and c[0, 22] == 1
    a = tc.canonical_alpha_dc(np.ones((3, 11)))
    assert a.shape == (3, 24) and np.all(a[:, 1] == 0) and np.all(a[:, 20] == tc.ORDER_DC_BOUND_DRIFT)


@pytest.mark.gpu
def test_gpu_fp64_shared_increments_reproduce_reference(sim, oracle):
    for name, canon, n, seed, ref in _cases():
        o = oracle.simulate_mt(7, canon, n, seed)
        normals = oracle.mt_normals(seed, int(o.n_steps.sum()) + 200 * n + 64)
        ob = oracle.simulate_buffer(7, canon, n, normals)
        off = np.zeros(n, np.int64)
        off[1:] = np.cumsum(ob.consumed)[:-1]
        sim.set_normals_debug(normals, off)
        try:
            out = sim.simulate(7, canon, n, precision=64, flags=F_STEPS, seed=1, dataset_offset=0)[0]
            steps = sim.last_steps(n)
            st = sim.last_stats()
        finally:
            sim.set_normals_debug(None, None)
        assert out.shape == (n, 3)
        got = np.ascontiguousarray(out[:, :ref.shape[1]])
        assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), name
        assert np.array_equal(steps, o.n_steps) and st["debug_overruns"] == 0


@pytest.mark.gpu
def test_gpu_production_kernel_consistency(sim, oracle):
    from bayesflow_nddms_b200 import two_channel as tc

    rng = np.random.default_rng(5)
    B, N = 64, 500
    P = np.column_stack([rng.normal(0, 2, B), rng.uniform(.6, 2, B), rng.uniform(.3, .7, B), rng.uniform(.1, .6, B),
                         rng.uniform(0, 2, B), rng.uniform(.5, 1.5, B), rng.uniform(0, 1, B), rng.normal(0, 1, B),
                         rng.normal(0, 1, B), rng.uniform(.1, 1, B), rng.uniform(.1, 1, B)])
    for canon in (tc.canonical_drift_dc5(P), tc.canonical_alpha_dc(P)):
        a = sim.simulate(7, canon, N, seed=9, dataset_offset=3, flags=F_STEPS)
        sa, st = sim.last_steps(B * N), sim.last_stats()
        assert a.shape == (B, N, 3) and st["used_persistent"] == 1 and st["reject_cap_hits"] == 0
        b = sim.simulate(7, canon, N, seed=9, dataset_offset=3, flags=F_STEPS | F_GENERIC)
        sb = sim.last_steps(B * N)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)) and np.array_equal(sa, sb)   # scheduling-invariant
        f = sim.simulate(7, canon, N, seed=9, dataset_offset=3, flags=F_F32)
        assert np.array_equal(f, a.astype(np.float32))
        # check #1 at scale: fp64 reference arithmetic on the same fp32 normals
        c = sim.simulate(7, canon, N, seed=9, dataset_offset=3, precision=64, flags=F_STEPS | 32)
        sc = sim.last_steps(B * N)
        same = sa == sc
        assert 1 - same.mean() < 2e-3
        a2, c2 = a.reshape(-1, 3), c.reshape(-1, 3)
        assert np.array_equal(a2[same, 0].view(np.uint64), c2[same, 0].view(np.uint64))
        close = np.abs(a2[same, 1:] - c2[same, 1:]) < 1e-4 * (1 + np.abs(c2[same, 1:]))
        assert close.mean() > 1 - 1e-3
        # sharding invariance
        lo = sim.simulate(7, canon[:20], N, seed=9, dataset_offset=3)
        hi = sim.simulate(7, canon[20:], N, seed=9, dataset_offset=23)
        assert np.array_equal(np.concatenate([lo, hi]), a)


@pytest.mark.gpu
@pytest.mark.parametrize("variant,params", [(0, [1.5, 1.2, 0.5, 0.4, 1.0, 1.0, 0.5, 0.7, -0.4, 0.3, 0.6]),
                                            (0, [-0.5, 0.9, 0.4, 0.3, 2.0, 0.3, 1.5, -1.2, 0.8, 0.9, 0.1]),
                                            (2, [2.0, 1.3, 0.55, 0.35, 0.4, 1.0, 0.5, 0.6, -0.3, 1.0, 2.0]),
                                            (2, [0.5, 0.3, 0.45, 0.3, 1.5, 0.4, 1.2, 1.1, 0.9, 0.2, 0.2])])
def test_gpu_distribution_vs_reference_loop(sim, oracle, variant, params):
    canon = _canon(variant, np.asarray(params, dtype=np.float64))
    n = 60_000
    g = sim.simulate(7, canon, n, seed=31, dataset_offset=2)[0]
    r = oracle.simulate_mt(7, canon, n, seed=77).sim_data
    for col in range(3):
        assert stats.ks_2samp(g[:, col], r[:, col]).pvalue > 1e-3, col
    # joint structure between the two channels and with |choicert|
    for i, j in ((1, 2), (0, 1), (0, 2)):
        cg = np.corrcoef(np.abs(g[:, i]) if i == 0 else g[:, i], g[:, j])[0, 1]
        cr = np.corrcoef(np.abs(r[:, i]) if i == 0 else r[:, i], r[:, j])[0, 1]
        assert abs(cg - cr) < 0.02, (i, j)


@pytest.mark.gpu
def test_gpu_two_channel_module_api(sim):
    import torch

    from bayesflow_nddms_b200 import two_channel as tc

    p = [1.5, 1.2, 0.5, 0.4, 1.0, 1.0, 0.5, 0.7, -0.4, 0.3, 0.6]
    for fn in (tc.simulate_trials_drift_dc5, tc.simulate_trials_drift_dc4, tc.simulate_trials_alpha_dc,
               tc.simulate_trials_drift_alpha):
        out = fn(p, 200, sim)
        assert out.shape == (200, 3) and out.dtype == np.float64 and np.all(np.isfinite(out))
    one = tc.simulate_trials_alpha_standardised([2.0, 1.2, 0.5, 0.35, 0.6, 1.0, 0.7], 5000, sim)
    assert one.shape == (5000, 2) and abs(one[:, 1].std() - 1) < 0.06
    d5 = tc.simulate_trials_drift_dc5(p, 20000, sim, seed=1, dataset_offset=0)
    assert abs(d5[:, 1].std() - 1) < 0.05 and abs(d5[:, 2].std() - 1) < 0.05     # standardised channels
    batch = tc.simulate_trials_alpha_dc(np.tile(p, (7, 1)), 64, sim)
    assert batch.shape == (7, 64, 3)
    sim.run(7, tc.canonical_drift_dc5(np.tile(p, (4, 1))), 33, flags=F_F32)
    t = torch.from_dlpack(sim.last_output_dlpack())
    assert tuple(t.shape) == (4, 33, 3) and t.dtype == torch.float32
