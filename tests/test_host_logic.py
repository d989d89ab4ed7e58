"""Host-side logic that needs no GPU: priors, configurator, sharding, Stahl preprocessing,
and the world_size-2 gather path on the gloo backend."""
import os
import socket
import sys

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT


# ---- priors: same distributions as the reference (basic_ddm_dc.py:62-80 etc.) -----------------
def test_prior_shapes_and_order():
    from bayesflow_nddms_b200 import basic_ddm_dc as m0
    from bayesflow_nddms_b200 import priors
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

    assert m0.draw_prior().shape == (5,) and m0.num_params == 5
    assert m1.draw_prior().shape == (7,) and m1.draw_prior_alt().shape == (7,) and m1.draw_prior_scale().shape == (8,)
    for name, cols in priors.PARAM_NAMES.items():
        assert priors.draw_prior_batch(name, 9, np.random.default_rng(0)).shape == (9, len(cols))
    n = [m0.prior_N() for _ in range(300)]
    assert min(n) >= 60 and max(n) <= 300


def test_prior_marginals_match_reference_distributions():
    from bayesflow_nddms_b200 import priors

    rng = np.random.default_rng(123)
    p = priors.draw_prior_batch("alpha_scale", 40000, rng)
    tn = lambda m, s, lo, hi: stats.truncnorm((lo - m) / s, (hi - m) / s, loc=m, scale=s)  # noqa: E731
    targets = [stats.norm(0, 2), tn(1, .5, 0, 10), stats.beta(2, 2), tn(.5, .25, 0, 1.5), tn(1, .5, 0, 3), tn(1, .5, 0, 10),
               stats.uniform(0, 5), stats.uniform(0, 2)]
    for j, d in enumerate(targets):
        assert stats.kstest(p[:, j], d.cdf).pvalue > 1e-3, j
    s = priors.draw_prior_batch("stahl", 40000, rng)
    for j, d in enumerate([stats.norm(3, 1), stats.beta(25, 25), tn(.4, .1, 0, 1.5), tn(1, .25, 0, 10)]):
        assert stats.kstest(s[:, j], d.cdf).pvalue > 1e-3, j
    sw = priors.draw_prior_batch("sweep", 100, rng)
    assert np.all(sw[:, 3] == 0)
    x = priors.truncnorm_better(mean=1.0, sd=0.5, low=0.0, upp=10)
    assert x.shape == (1,) and 0 <= x[0] <= 10


def test_reference_prior_agrees_when_available():
    """In the build container the verbatim reference prior can be exec'd: compare marginals."""
    from oracle import ref_loader as rl

    if not rl.available():
        pytest.skip("reference tree not present (GPU box)")
    from bayesflow_nddms_b200 import priors

    ns = rl.load("basic_prior")
    np.random.seed(12345)  # the reference's truncnorm draws use NumPy's global state: fix it, no flaky KS
    ref = np.stack([ns["draw_prior"]() for _ in range(1500)])
    mine = priors.draw_prior_batch("basic", 20000, np.random.default_rng(5))
    for j in range(5):
        assert stats.ks_2samp(ref[:, j], mine[:, j]).pvalue > 1e-3, j


# ---- configurator ---------------------------------------------------------------------------------
def test_configurator_contract():
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

    d = {'sim_data': np.random.default_rng(0).normal(size=(4, 77, 2)), 'sim_non_batchable_context': 77,
         'prior_draws': np.ones((4, 7))}
    c = m1.configurator(d)
    assert c['summary_conditions'].dtype == np.float32 and c['summary_conditions'].shape == (4, 77, 2)
    assert c['direct_conditions'].shape == (4, 1) and np.allclose(c['direct_conditions'], np.log(77))
    assert c['parameters'].dtype == np.float32
    d['prior_draws'] = None                  # single_trial_alpha_not_scaled.py:189
    assert 'parameters' not in m1.configurator(d)


def test_device_configurator_on_cpu_tensor():
    import torch

    from bayesflow_nddms_b200 import basic_ddm_dc as m0

    d = {'sim_data': torch.ones(3, 10, 2, dtype=torch.float64), 'sim_non_batchable_context': 10, 'prior_draws': np.zeros((3, 5))}
    c = m0.device_configurator(d)
    assert c['summary_conditions'].dtype == torch.float32 and c['direct_conditions'].shape == (3, 1)
    assert abs(float(c['direct_conditions'][0, 0]) - np.log(10)) < 1e-6


# ---- Stahl preprocessing (imputation_from_stahl_not_scaled.py:82-105) ---------------------------------
def test_stahl_boundaries():
    from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as st

    subj, pe = st.synthetic_stahl_like()
    like, alphas = st.boundaries_from_pe(pe)
    z = (pe - pe.mean()) / pe.std()
    assert np.array_equal(like, (z + 3) / 3)
    assert np.array_equal(alphas, np.maximum((z + 3) / 3, 0) * ((z + 3) / 3 >= 0))
    assert alphas.min() == 0 and like.min() < 0
    counts = np.unique(subj, return_counts=True)[1]
    assert counts.size == 89 and counts.min() >= 13 and counts.sum() == 19374


def test_stahl_csv_when_available():
    from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as st

    path = "/root/reference/stahl_data/base_data.csv"
    if not os.path.exists(path):
        pytest.skip("reference data not present (GPU box)")
    subj, pe = st.load_stahl_csv(path)
    assert subj.size == 19374 and np.unique(subj).size == 89
    like, alphas = st.boundaries_from_pe(pe)
    assert alphas.min() >= 0 and abs(like.mean() - 1.0) < 1e-9


# ---- sharding -----------------------------------------------------------------------------------------
def test_shard_range_partitions():
    from bayesflow_nddms_b200.distributed import shard_range

    for n in (0, 1, 7, 64, 1000, 1_000_001):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from bayesflow_nddms_b200 import distributed as D
    from oracle import cpu as orc

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    params = np.tile(np.array([[1.0, 1.2, 0.5, 0.3, 1.0]]), (n_total, 1))
    params[:, 0] = np.linspace(-2, 2, n_total)

    def fake_simulate(p, n_trials, dataset_offset=0):
        # stand-in for the CUDA call with the same keying: the global dataset index picks the stream
        return np.stack([orc.simulate_philox(0, p[i], n_trials, 99, dataset=dataset_offset + i).sim_data
                         for i in range(p.shape[0])]) if p.shape[0] else np.empty((0, n_trials, 2))

    full, (lo, hi) = D.simulate_sharded(fake_simulate, params, 20, gather=True)           # batch counter: base 0
    nxt, _ = D.simulate_sharded(fake_simulate, params, 20, gather=True)                   # ... advanced by n_total
    local, _ = D.simulate_sharded(fake_simulate, params, 20, gather=False, dataset_base=0)  # an explicit base regenerates
    q.put((rank, lo, hi, full, local, nxt))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_sharded_simulation_with_gloo_gather_world2(n_total, oracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params = np.tile(np.array([[1.0, 1.2, 0.5, 0.3, 1.0]]), (n_total, 1))
    params[:, 0] = np.linspace(-2, 2, n_total)
    single = np.stack([oracle.simulate_philox(0, params[i], 20, 99, dataset=i).sim_data for i in range(n_total)])
    (r0, lo0, hi0, full0, loc0, nxt0), (r1, lo1, hi1, full1, loc1, nxt1) = res
    assert (lo0, hi1) == (0, n_total) and hi0 == lo1
    assert np.array_equal(full0, single) and np.array_equal(full1, single)   # independent of world size
    assert np.array_equal(np.concatenate([loc0, loc1]), single)
    # successive batches use fresh counters: the second one is keyed n_total datasets further on, on every rank
    second = np.stack([oracle.simulate_philox(0, params[i], 20, 99, dataset=n_total + i).sim_data for i in range(n_total)])
    assert np.array_equal(nxt0, second) and np.array_equal(nxt1, second) and not np.array_equal(second, single)


# ---- alpha_not_scaled.py data generation (:52-131): participant parameters are the reference's own ---------
def test_alpha_not_scaled_participant_parameters_match_reference():
    from bayesflow_nddms_b200 import alpha_not_scaled as m

    z = np.load(os.path.join(ROOT, "tests", "golden", "alpha_not_scaled_params.npz"))  # lines 54-88 run verbatim
    g, _ = m.draw_participants(100, 2, 2021)
    for k in ("ndt", "alpha", "beta", "delta", "varsigma", "deltatrialsd"):
        assert np.array_equal(g[k], z[k]), k
    assert g["sigma"] == float(z["sigma"]) and g["var_alpha"] == float(z["var_alpha"])
    assert (g["ndt"][17], g["alpha"][17], g["delta"][17], g["deltatrialsd"][17]) == (.4, 1.2, 3.5, 1)
    assert m.draw_participants(100, 4)[0]["sigma"] == .2


def test_device_replay_buffer_fifo_and_sampling():
    import torch

    from bayesflow_nddms_b200.replay import DeviceReplayBuffer

    buf = DeviceReplayBuffer(3, rng=np.random.default_rng(0))
    with pytest.raises(RuntimeError):
        buf.sample()
    for i in range(5):
        buf.store({'summary_conditions': torch.full((2, 4, 2), float(i)), 'parameters': torch.zeros(2, 5)})
    assert len(buf) == 3 and buf.is_full() and buf.stored_total == 5
    seen = {float(buf.sample()['summary_conditions'][0, 0, 0]) for _ in range(200)}
    assert seen == {2.0, 3.0, 4.0}                       # the two oldest batches were overwritten
    assert buf.nbytes() == 3 * (2 * 4 * 2 + 2 * 5) * 4


@pytest.mark.parametrize("basic", [True, False])
@pytest.mark.parametrize("f32", [False, True])
def test_wire_decode_host_formats_rows_like_the_reference(basic, f32):
    """Host half of the compact device->host format (include/ddm_b200.h: ddm_wire_decode_host): from the
    kernel's integers it must produce basic_ddm_dc.py:108-112's rt = n*dt + ndt and choice (and the signed-rt /
    external-measurement rows of single_trial_alpha_not_scaled.py:131-141), whatever the thread count."""
    import ctypes as C

    from bayesflow_nddms_b200 import _capi

    lib = _capi.load()
    rng = np.random.default_rng(3)
    D, T, dt = 37, 101, 0.01
    params = rng.uniform(0.1, 1.0, (D, 6))
    n = rng.integers(0, 401, (D, T)).astype(np.uint32)
    ch = rng.integers(-1, 2, (D, T))
    code = ((n << 2) | (ch + 1).astype(np.uint32)).astype(np.int32)
    ext = rng.standard_normal((D, T)).astype(np.float32)
    if basic:
        wire = code
    else:
        wire = np.stack([code, ext.view(np.int32)], axis=-1).copy()
    rt = n.astype(np.float64) * dt
    tau = params[:, 3:4]
    for flags in (0, 1):
        if basic:
            want = np.stack([rt + tau, np.where(ch == 0, float(flags & 1), ch).astype(np.float64)], axis=-1)
        else:
            want = np.stack([np.where(ch > 0, tau + rt, np.where(ch < 0, -tau - rt, 0.0)), ext.astype(np.float64)], axis=-1)
        if f32:
            want = want.astype(np.float32)
        for threads in (1, 2, 5):
            out = np.full((D, T, 2), np.nan, dtype=np.float32 if f32 else np.float64)
            rc = lib.ddm_wire_decode_host(wire.ctypes.data, out.ctypes.data, params.ctypes.data_as(_capi._dp), 6, D, T, dt,
                                          int(basic), flags | (_capi.FLAG_OUT_F32 if f32 else 0), threads)
            assert rc == 0
            assert np.array_equal(out, want), (flags, threads)
    # misaligned destination: falls back to ordinary stores
    if not f32:
        raw = np.empty(D * T * 2 + 1)
        out = raw[1:].reshape(D, T, 2)
        assert lib.ddm_wire_decode_host(wire.ctypes.data, out.ctypes.data, params.ctypes.data_as(_capi._dp), 6, D, T, dt,
                                        int(basic), 1, 3) == 0
        assert np.array_equal(out, want)


def test_wire_decode_float32_streaming_path_any_alignment():
    """Round 2: float32 rows are written with 32-byte streaming stores, eight trials at a time (ddm_wire.cpp:
    decode_basic32_avx2).  Destinations at every 4-byte phase of a 32-byte line, lengths around the vector width,
    per-dataset non-decision times: always (float)(n*dt + tau) rounded once from double, like the kernel's own rows."""
    from bayesflow_nddms_b200 import _capi

    lib = _capi.load()
    rng = np.random.default_rng(8)
    dt = 1e-3
    for D, T in ((3, 1000), (5, 17), (1, 7), (2, 8), (1, 40_003)):
        params = rng.uniform(0.0, 1.5, (D, 5))
        n = rng.integers(0, 4001, (D, T)).astype(np.uint32)
        ch = rng.integers(-1, 2, (D, T))
        code = ((n << 2) | (ch + 1).astype(np.uint32)).astype(np.int32)
        want = np.stack([(n.astype(np.float64) * dt + params[:, 3:4]).astype(np.float32), ch.astype(np.float32)], axis=-1)
        for phase in range(8):
            raw = np.full(D * T * 2 + 16, np.nan, dtype=np.float32)
            base = (-raw.ctypes.data // 4) % 8             # first 32-byte aligned element
            out = raw[base + phase: base + phase + D * T * 2].reshape(D, T, 2)
            for threads in (1, 3):
                out[...] = np.nan
                assert lib.ddm_wire_decode_host(code.ctypes.data, out.ctypes.data, params.ctypes.data_as(_capi._dp), 5, D, T, dt, 1,
                                                _capi.FLAG_OUT_F32, threads) == 0
                assert np.array_equal(out, want), (D, T, phase, threads)
            assert np.isnan(raw[:base + phase]).all() and np.isnan(raw[base + phase + D * T * 2:]).all()   # nothing outside the rows


@pytest.mark.parametrize("f32", [False, True])
def test_wire_decode_external_column_layout_streaming_path_any_alignment(f32):
    """The (signed rt, external measurement) rows of the single-trial-boundary models through the vectorised decoders
    (decode_ext64_avx2 / decode_ext32_avx2): every phase of the destination, lengths around the vector width, negative
    and zero non-decision times (a missing response must stay +0.0), all three choices."""
    from bayesflow_nddms_b200 import _capi

    lib = _capi.load()
    rng = np.random.default_rng(9)
    dt = 0.01
    dtype = np.float32 if f32 else np.float64
    for D, T in ((3, 1000), (5, 17), (1, 3), (2, 8), (1, 20_003)):
        params = rng.uniform(-0.5, 1.5, (D, 7))
        params[0, 3] = 0.0
        n = rng.integers(0, 401, (D, T)).astype(np.uint32)
        ch = rng.integers(-1, 2, (D, T))
        code = ((n << 2) | (ch + 1).astype(np.uint32)).astype(np.int32)
        ext = rng.standard_normal((D, T)).astype(np.float32)
        wire = np.stack([code, ext.view(np.int32)], axis=-1).copy()
        rt, tau = n.astype(np.float64) * dt, params[:, 3:4]
        want = np.stack([np.where(ch > 0, tau + rt, np.where(ch < 0, -tau - rt, 0.0)), ext.astype(np.float64)], axis=-1).astype(dtype)
        per = 32 // dtype().itemsize
        for phase in range(per):
            raw = np.full(D * T * 2 + 2 * per, np.nan, dtype=dtype)
            base = (-raw.ctypes.data // dtype().itemsize) % per
            out = raw[base + phase: base + phase + D * T * 2].reshape(D, T, 2)
            for threads in (1, 3):
                out[...] = np.nan
                assert lib.ddm_wire_decode_host(wire.ctypes.data, out.ctypes.data, params.ctypes.data_as(_capi._dp), 7, D, T, dt, 0,
                                                _capi.FLAG_OUT_F32 if f32 else 0, threads) == 0
                assert np.array_equal(out, want) and not np.signbit(out[..., 0][ch == 0]).any(), (D, T, phase, threads)
            assert np.isnan(raw[:base + phase]).all() and np.isnan(raw[base + phase + D * T * 2:]).all()


def test_host_stream_store_peak_is_measurable():
    import ctypes as C

    from bayesflow_nddms_b200 import _capi

    lib = _capi.load()
    v = C.c_double()
    assert lib.ddm_host_stream_peak(2, 64 << 20, C.byref(v)) == 0 and v.value > 1e8      # > 0.1 GB/s on anything
    assert lib.ddm_host_stream_peak(2, 1000, C.byref(v)) == _capi.ERR_INVALID


def test_simulator_counters_roll_over_multiples_of_2_32():
    """The 64-bit global index: a launch must not straddle a multiple of 2^32, the Python counters skip
    ahead instead of failing; distributed.simulate_sharded's batch counter does the same."""
    from bayesflow_nddms_b200 import distributed as D
    from bayesflow_nddms_b200.simulator import DDMSimulator

    roll = DDMSimulator._roll
    assert roll(10, 5) == 10 and roll((1 << 32) - 5, 5) == (1 << 32) - 5 and roll((1 << 32) - 4, 5) == 1 << 32
    assert roll((7 << 32) + 12, 1 << 32) == 8 << 32 and roll(0, 1 << 32) == 0
    seen = []
    D._batch_base = (1 << 32) - 3
    try:
        for _ in range(2):
            D.simulate_sharded(lambda p, n, dataset_offset: seen.append(dataset_offset) or np.zeros((p.shape[0], n, 2)),
                               np.zeros((8, 5)), 4, rank=0, world_size=1)
    finally:
        D._batch_base = 0
    assert seen == [1 << 32, (1 << 32) + 8]


@pytest.mark.parametrize("n_datasets,n_trials", [(1_000_000, 1000), (70_000, 1000), (4_200, 1000), (7, 30_000_000), (3, 5), (0, 10),
                                                 (100_003, 777)])
@pytest.mark.parametrize("min_chunk_rows", [-1, 1, 50_000, 4 << 20])
def test_streamed_histogram_chunk_schedule(n_datasets, n_trials, min_chunk_rows):
    """include/ddm_b200.h: ddm_histogram_chunks -- every dataset exactly once, in order; a sixteenth of the batch first
    (the GPU starts after a small upload), then half of what is left each time, no chunk below the minimum."""
    import ctypes as C

    from bayesflow_nddms_b200 import _capi

    lib = _capi.load()
    cap = 4096
    first = (C.c_int64 * cap)()
    count = (C.c_int64 * cap)()
    n = lib.ddm_histogram_chunks(n_datasets, n_trials, min_chunk_rows, first, count, cap)
    if n_datasets == 0:
        assert n == 0
        return
    assert 1 <= n <= 64                                    # geometric: a handful of launches whatever the size
    f, c = np.array(first[:n]), np.array(count[:n])
    assert f[0] == 0 and np.all(c >= 1) and np.array_equal(f[1:], np.cumsum(c)[:-1]) and f[-1] + c[-1] == n_datasets
    floor_rows = (32 << 20) if min_chunk_rows <= 0 else min_chunk_rows
    floor = max(1, -(-floor_rows // n_trials))             # in datasets
    assert np.all(c >= min(floor, n_datasets))
    if n > 1:
        assert c[0] == max(n_datasets // 16, floor)
        left = n_datasets - f[1:]
        assert np.all((c[1:] == np.maximum(left // 2, floor)) | (c[1:] == left))   # half of what is left, or the rest
    assert lib.ddm_histogram_chunks(-1, 5, 0, None, None, 0) == -1


@pytest.mark.parametrize("n_datasets,n_trials", [(1_000_000, 1000), (16_384, 1000), (4_200, 1000), (7, 3_000_000), (3, 5), (0, 10),
                                                 (100_003, 33)])
@pytest.mark.parametrize("chunk_rows", [-1, -2, -3, 1, 257 * 7, 32 << 20])
def test_streamed_path_chunk_schedule(n_datasets, n_trials, chunk_rows):
    """include/ddm_b200.h: ddm_pipeline_chunks -- the schedule covers every dataset exactly once, in order; the
    default schedules keep chunks within 2 Mi .. 32 Mi trials (512 Ki for batches below 4 Mi; whole datasets) and end with
    small chunks."""
    import ctypes as C

    from bayesflow_nddms_b200 import _capi

    lib = _capi.load()
    cap = 1 << 16
    first = (C.c_int64 * cap)()
    count = (C.c_int64 * cap)()
    n = lib.ddm_pipeline_chunks(n_datasets, n_trials, chunk_rows, first, count, cap)
    if n_datasets == 0:
        assert n == 0
        return
    assert 1 <= n
    if n > cap:                                   # tiny fixed chunks of a large batch: count only
        per = max(1, chunk_rows // n_trials)
        assert chunk_rows > 0 and n == -(-n_datasets // per)
        return
    f, c = np.array(first[:n]), np.array(count[:n])
    assert f[0] == 0 and np.all(c >= 1) and np.array_equal(f[1:], np.cumsum(c)[:-1]) and f[-1] + c[-1] == n_datasets
    rows = c * n_trials
    if chunk_rows < 0:
        assert np.all(rows <= max(32 << 20, n_trials))                 # at most 32 Mi trials, or one dataset
        if n > 1:
            floor = (512 << 10) if n_datasets * n_trials < (4 << 20) else (2 << 20)   # small batches: 512 Ki-trial chunks
            assert np.all(rows[:-1] >= min(floor, rows[:-1].max()) - n_trials)        # ... up to dataset granularity
        if n_datasets * n_trials >= 256 << 20 and chunk_rows != -2:
            assert rows[-1] <= 4 << 20 < rows[0]                        # large batches end with small chunks
    elif chunk_rows > 0:
        assert np.all(c[:-1] == max(1, chunk_rows // n_trials))


def test_unique_inverse_is_np_unique():
    """imputation_from_stahl_not_scaled.unique_inverse: the participant index without a sort, same arrays as
    np.unique(..., return_inverse=True) for every kind of id column."""
    from bayesflow_nddms_b200.imputation_from_stahl_not_scaled import synthetic_stahl_like, unique_inverse

    rng = np.random.default_rng(5)
    subj, _ = synthetic_stahl_like()
    cases = [subj, subj.astype(np.int32), rng.integers(-50, 50, 1000), rng.integers(0, 3, 10).astype(np.uint8),
             np.array([7]), np.array([], dtype=np.int64), rng.integers(0, 10**12, 500),      # sparse: falls back to the sort
             rng.normal(size=100), np.array(["b", "a", "b"])]
    for a in cases:
        u1, i1 = np.unique(a, return_inverse=True)
        u2, i2 = unique_inverse(a)
        assert u1.dtype == u2.dtype and np.array_equal(u1, u2) and np.array_equal(i1, i2), a[:5]


def test_boundaries_from_pe_is_the_reference_arithmetic():
    """imputation_from_stahl_not_scaled.py:82-105 written out literally gives the bits boundaries_from_pe returns
    (which forms the centred column once and copies the two identical columns)."""
    from bayesflow_nddms_b200.imputation_from_stahl_not_scaled import boundaries_from_pe

    rng = np.random.default_rng(9)
    for n in (2, 3, 13, 337, 19374, 100_003):
        for scale in (1.0, 1e-3, 1e3):
            all_Pe = np.clip(rng.normal(0.0, 5.8, n), -35, 35) * scale
            all_standard_Pe = (all_Pe - np.mean(all_Pe)) / np.std(all_Pe)
            alpha_like_Pe = (all_standard_Pe + 3) / 3
            single_trial_alphas = (all_standard_Pe + 3) / 3
            single_trial_alphas[single_trial_alphas < 0] = 0
            a, b = boundaries_from_pe(all_Pe)
            assert np.array_equal(a, alpha_like_Pe) and np.array_equal(b, single_trial_alphas) and b.min() >= 0
            assert a is not b and not np.shares_memory(a, b)
