"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
and the committed reference outputs.  Run on a B200 with `pytest -m gpu`."""
import numpy as np
import pytest

from conftest import VARIANT_MODEL, variant_kwargs

pytestmark = pytest.mark.gpu

F_TIMEOUT1, F_F32, F_STEPS, F_GENERIC, F_STATE = 1, 2, 4, 8, 16


def _offsets(consumed):
    off = np.zeros(consumed.size, np.int64)
    off[1:] = np.cumsum(consumed)[:-1]
    return off


def _trial_increments(sim, model, params, ds, t, n_step_normals, seed):
    """The production kernel's own normals of one trial in the reference's consumption order
    (pre-draws, ``n_step_normals`` step normals, ext-data draw): north_star's exported increments."""
    zs = sim.export_normals(ds, t, 0, 0, int(n_step_normals), seed=seed) if n_step_normals else np.empty(0)
    if model in (0, 5):
        return zs
    if model == 6:
        return np.concatenate([sim.export_normals(ds, t, 1, 1, 1, seed=seed), zs])
    aux = sim.export_normals(ds, t, 1, 0, 4097, seed=seed)
    mu, sd = (params[5], params[4]) if model == 2 else (params[1], params[4])
    cand = np.float32(mu) + np.float32(sd) * aux[1:].astype(np.float32)
    k = int(np.argmax(cand > 0))          # index of the first positive candidate
    return np.concatenate([aux[1:k + 2], zs, aux[:1]])


def _reference_trial(sim, oracle, model, params, ds, t, n_gpu, seed, kw, bound_in=None):
    """The reference's fp64 loop (CPU oracle) on the kernel's exported increments of one trial.  The
    buffer first holds exactly the kernel's step count of step normals (so that the draw made after the
    loop lands where the reference draws it); if the fp64 path wants more steps it is rerun with more."""
    for extra in (0, 64, int(kw["max_steps"])):
        S = min(int(n_gpu) + extra, int(kw["max_steps"]))
        z = _trial_increments(sim, model, params, ds, t, S, seed)
        try:
            r = oracle.simulate_buffer(model, params, 1, z, bound_in=bound_in, **kw)
        except IndexError:
            continue
        if int(r.n_steps[0]) <= S:
            return r, z
    raise AssertionError(f"trial {t}: the reference loop did not finish on {S} step normals")


ULP32 = 2.0 ** -23
TIE_ULPS = 64   # a tie = the fp64 path is within this many fp32 ulps (of the boundary scale) of a boundary


def _assert_boundary_tie(sim, oracle, model, params, ds, t, n_gpu, n_ref, seed, kw, bound_in=None):
    """SURVEY 7 / VERDICT r1: a trial whose crossing step differs between the fp32 kernel and the fp64
    reference loop on the same increments is a *documented boundary tie* only if, at the step where the
    two paths part (the earlier of the two crossing steps), the fp64 evidence sits within a few fp32 ulps
    of the boundary it is compared with.  Returns that distance in fp32 ulps of the boundary scale."""
    m = int(min(n_gpu, n_ref))
    what = f"model {model} dataset {ds} trial {t}"
    assert m >= 1, f"{what}: paths part before the first step"
    z = _trial_increments(sim, model, params, ds, t, m, seed)
    r = oracle.simulate_buffer(model, params, 1, z, bound_in=bound_in, **dict(kw, max_steps=float(m)))
    assert int(r.n_steps[0]) == m, what
    ev, bound = float(r.evidence[0]), float(r.bound[0])
    dist = min(abs(ev), abs(ev - bound))
    ulps = dist / (ULP32 * max(bound, abs(ev), 1e-30))
    assert ulps <= TIE_ULPS, (f"{what}: crossing steps {n_gpu} (fp32 kernel) vs {n_ref} (fp64 loop) but the fp64 evidence "
                              f"{ev!r} at step {m} is {ulps:.1f} fp32 ulps from a boundary (0, {bound!r}): not a tie")
    return ulps


# --------------------------------------------------------------------------------------------
# Philox and the normal map
# --------------------------------------------------------------------------------------------
def test_philox_known_answers_on_device(sim, oracle):
    ctr = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32)
    key = np.array([[0, 0], [0xffffffff] * 2, [0xa4093822, 0x299f31d0]], np.uint32)
    got = sim.philox4x32(ctr, key)
    want = np.array([[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8],
                     [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd],
                     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]], np.uint32)
    assert np.array_equal(got, want)
    rng = np.random.default_rng(5)
    ctr = rng.integers(0, 2**32, (257, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2**32, (257, 2), dtype=np.uint64).astype(np.uint32)
    got = sim.philox4x32(ctr, key)
    want = np.stack([oracle.philox4x32_10(c, k) for c, k in zip(ctr, key)])
    assert np.array_equal(got, want)


def test_exported_normals_match_ideal_map(sim, oracle):
    seed = 0x1234567890abcdef
    for stream in (0, 1):
        ideal = oracle.philox_normals(seed, 11, 77, stream, 3, 2001)
        z64 = sim.export_normals(11, 77, stream, 3, 2001, seed=seed, precision=64)
        z32 = sim.export_normals(11, 77, stream, 3, 2001, seed=seed, precision=32)
        # fp64 map: libdevice log/sincospi vs glibc -- a few ulp
        assert np.max(np.abs(z64 - ideal)) < 1e-13
        # fp32 production map: MUFU lg2/sqrt/sin/cos approximations
        assert np.max(np.abs(z32 - ideal)) < 2e-5
        assert np.all(z32 == z32.astype(np.float32))  # exported values are the fp32 numbers


# --------------------------------------------------------------------------------------------
# Check #1a: fp64 kernel on shared increments == the reference's own output, bit for bit
# --------------------------------------------------------------------------------------------
def test_fp64_shared_increments_reproduce_reference_outputs(sim, oracle, golden):
    z, meta = golden
    for name, m in meta.items():
        if name == "stahl":
            continue
        model = VARIANT_MODEL[m["variant"]]
        kw = variant_kwargs(m["variant"])
        params, n = z[f"{name}__params"], m["n_trials"]
        flags = F_TIMEOUT1 if model in (0, 6) else 0
        o = oracle.simulate_mt(model, params, n, m["seed"], flags=flags, **kw)
        total = int(o.n_steps.sum()) + 64 * n + 64
        normals = oracle.mt_normals(m["seed"], total)
        ob = oracle.simulate_buffer(model, params, n, normals, flags=flags, **kw)
        sim.set_normals_debug(normals, _offsets(ob.consumed))
        try:
            out = sim.simulate(model, params, n, precision=64, flags=flags | F_STEPS, seed=1, dataset_offset=0, **kw)
            steps = sim.last_steps(n)
            st = sim.last_stats()
            ev = sim.simulate(model, params, n, precision=64, flags=flags | F_STATE, seed=1, dataset_offset=0, **kw)
        finally:
            sim.set_normals_debug(None, None)
        ref = z[f"{name}__out"]
        assert np.array_equal(out[0].view(np.uint64), ref.view(np.uint64)), name
        assert np.array_equal(steps, o.n_steps), name
        assert np.array_equal(ev[0, :, 1].view(np.uint64), o.evidence.view(np.uint64)), name
        assert st["debug_overruns"] == 0 and st["used_persistent"] == 0
        assert st["total_steps"] == int(o.n_steps.sum())
        assert st["n_timeouts"] == int((o.choice == 0).sum())


def test_fp64_shared_increments_stahl(sim, oracle, golden):
    z, meta = golden
    m = meta["stahl"]
    bounds, p = z["stahl__bounds"], z["stahl__params"]
    o = oracle.simulate_mt(5, p, m["n_trials"], m["seed"], bound_in=bounds)
    normals = oracle.mt_normals(m["seed"], int(o.n_steps.sum()) + 64)
    ob = oracle.simulate_buffer(5, p, m["n_trials"], normals, bound_in=bounds)
    sim.set_normals_debug(normals, _offsets(ob.consumed))
    try:
        out = sim.simulate_trialwise(np.zeros(bounds.size, np.int32), bounds, p[None, :], precision=64)
    finally:
        sim.set_normals_debug(None, None)
    assert np.array_equal(out[:, 0].view(np.uint64), z["stahl__out"].view(np.uint64))
    assert np.array_equal(out[:, 1], bounds)


# --------------------------------------------------------------------------------------------
# Check #1b: fp32 production kernel vs the reference fp64 loop on the kernel's own exported
# increments: choices and crossing steps bit-exact except boundary ties, states within 1e-5
# --------------------------------------------------------------------------------------------
PROD_CASES = [
    (0, [3.0, 1.5, 0.5, 0.4, 1.0], dict(dt=0.01, max_steps=400)),
    (0, [-0.8, 1.1, 0.35, 0.3, 0.7], dict(dt=0.001, max_steps=4000)),
    (0, [0.05, 4.0, 0.5, 0.3, 0.3], dict(dt=0.01, max_steps=400)),        # mostly timeouts
    (1, [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1], dict(dt=0.01, max_steps=400)),
    (1, [0.5, 0.2, 0.4, 0.3, 2.5, 1.2, 3.0], dict(dt=0.01, max_steps=400)),  # many boundary redraws
    (2, [-1.0, 1.0, 0.55, 0.2, 2.0, 0.3, 2.0], dict(dt=0.01, max_steps=400)),
    (3, [3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1, 1.37], dict(dt=0.01, max_steps=400)),
    (4, [1.0, 1.5, 0.6, 0.4, 0.5, 1.0, 0.1], dict(dt=0.001, max_steps=4000)),
    (6, [3.5, 1.2, 0.5, 0.4, 1.0, 1.2], dict(dt=0.01, max_steps=400)),       # per-trial drift (eta)
    (6, [-0.5, 1.4, 0.4, 0.3, 2.0, 0.9], dict(dt=0.001, max_steps=4000)),
]


@pytest.mark.parametrize("model,params,kw", PROD_CASES)
def test_fp32_production_vs_reference_loop_on_exported_increments(sim, oracle, model, params, kw):
    n, seed, ds = 384, 99, 5
    out = sim.simulate(model, params, n, precision=32, flags=F_STEPS, seed=seed, dataset_offset=ds, **kw)[0]
    steps = sim.last_steps(n)
    assert sim.last_stats()["used_persistent"] == 1
    state = sim.simulate(model, params, n, precision=32, flags=F_STATE, seed=seed, dataset_offset=ds, **kw)[0, :, 1]
    ref_steps = np.empty(n, np.int64)
    ref_choice = np.empty(n, np.int32)
    ref_ev = np.empty(n)
    ref_out = np.empty((n, 2))
    ref_bound = np.empty(n)
    for t in range(n):  # the reference loop on this trial's exported increments (pre-draws, steps, ext)
        r, _ = _reference_trial(sim, oracle, model, params, ds, t, steps[t], seed, kw)
        ref_steps[t], ref_choice[t], ref_ev[t], ref_out[t], ref_bound[t] = (
            r.n_steps[0], r.choice[0], r.evidence[0], r.sim_data[0], r.bound[0])
    same = (steps == ref_steps)
    # documented boundary ties: fp32 rounding flips a strict comparison only when the fp64 path passes
    # within rounding distance of a boundary -- at most one of the 384 trials, and it must *be* such a tie
    assert (~same).sum() <= 1, f"{(~same).sum()} of {n} crossing steps differ"
    for t in np.flatnonzero(~same):
        _assert_boundary_tie(sim, oracle, model, params, ds, t, steps[t], ref_steps[t], seed, kw)
    rt_col = out[:, 0]
    if model in (0, 6):
        gpu_choice = out[:, 1].astype(np.int32)
    else:
        gpu_choice = np.sign(rt_col).astype(np.int32)
    assert np.array_equal(gpu_choice[same], ref_choice[same])
    # identical step count => identical reported RT bits (fp64 output arithmetic of the reference)
    assert np.array_equal(rt_col[same].view(np.uint64), ref_out[same, 0].view(np.uint64))
    scale = np.maximum(ref_bound, 1e-3)
    assert np.max(np.abs(state[same] - ref_ev[same]) / scale[same]) < 1e-5
    if model not in (0, 6):
        assert np.max(np.abs(out[same, 1] - ref_out[same, 1])) < 1e-5 * (1 + np.max(np.abs(ref_out[:, 1])))


# --------------------------------------------------------------------------------------------
# Scheduling cannot change results
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", [0, 1, 2, 3, 4, 6])
def test_persistent_equals_generic_bitwise(sim, model):
    rng = np.random.default_rng(model)
    from bayesflow_nddms_b200 import priors

    name = ["basic", "alpha", "alpha_dc", "alpha_scale", "alpha_scale2", None, "eta"][model]
    params = priors.draw_prior_batch(name, 37, rng)
    for kw in (dict(dt=0.01, max_steps=400), dict(dt=0.001, max_steps=4000)):
        a = sim.simulate(model, params, 211, precision=32, flags=F_STEPS, seed=3, dataset_offset=10, **kw)
        sa, st = sim.last_steps(37 * 211), sim.last_stats()
        assert st["used_persistent"] == 1
        b = sim.simulate(model, params, 211, precision=32, flags=F_STEPS | F_GENERIC, seed=3, dataset_offset=10, **kw)
        sb, st2 = sim.last_steps(37 * 211), sim.last_stats()
        assert st2["used_persistent"] == 0
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
        assert np.array_equal(sa, sb)
        for k in ("total_steps", "n_timeouts", "n_upper", "reject_cap_hits"):
            assert st[k] == st2[k]
        assert st["total_steps"] == int(sa.sum())


def test_results_independent_of_tuning_and_sharding(sim):
    from bayesflow_nddms_b200 import priors

    params = priors.draw_prior_batch("basic", 64, np.random.default_rng(1))
    base = sim.simulate(0, params, 500, seed=11, dataset_offset=1000)
    try:
        for thr, bps, tile in [(1, 1, 7), (32, 2, 500), (16, 0, 32), (4, 3, 1)]:
            sim.set_tuning(thr, bps, tile)
            again = sim.simulate(0, params, 500, seed=11, dataset_offset=1000)
            assert np.array_equal(base, again), (thr, bps, tile)
    finally:
        sim.set_tuning(0, 0, 0)
    # two "ranks": the global dataset index keys the stream, not the launch
    lo = sim.simulate(0, params[:40], 500, seed=11, dataset_offset=1000)
    hi = sim.simulate(0, params[40:], 500, seed=11, dataset_offset=1040)
    assert np.array_equal(base, np.concatenate([lo, hi]))
    other = sim.simulate(0, params, 500, seed=12, dataset_offset=1000)
    assert not np.array_equal(base, other)


def test_pipelined_host_transfer_equals_single_launch(sim):
    """Large host-destined batches are simulated chunk by chunk while the previous chunk is copied
    back; the chunking must not change a bit, for pageable and for pinned destinations."""
    from bayesflow_nddms_b200 import priors

    params = priors.draw_prior_batch("alpha", 301, np.random.default_rng(4))
    sim.set_pipeline(1 << 60, -1)
    base = sim.simulate(1, params, 257, seed=6, dataset_offset=77)
    st0 = sim.last_stats()
    try:
        for chunk_rows in (257 * 7, 257 * 100, 257 * 300, 1):
            sim.set_pipeline(1, chunk_rows)
            again = sim.simulate(1, params, 257, seed=6, dataset_offset=77)
            st = sim.last_stats()
            assert np.array_equal(base, again), chunk_rows
            for k in ("total_steps", "n_timeouts", "n_upper", "n_trials"):
                assert st[k] == st0[k], (k, chunk_rows)
            with pytest.raises(Exception):
                sim.last_output_dlpack()        # the batch was streamed to the host, nothing is resident
        pinned = sim.pinned_empty((301, 257, 2), np.float64)
        sim.set_pipeline(1, 257 * 50)
        out = sim.simulate(1, params, 257, seed=6, dataset_offset=77, out=pinned)
        assert out is pinned and np.array_equal(base, out)
        f32 = sim.simulate(1, params, 257, seed=6, dataset_offset=77, flags=F_F32)
        assert np.array_equal(f32, base.astype(np.float32))
    finally:
        sim.set_pipeline(-1, -1)
    # default settings: small batches take the single-launch path and stay resident
    sim.simulate(1, params, 257, seed=6, dataset_offset=77)
    assert sim.last_output_device_ptr()[1] == 301 * 257 * 16


@pytest.mark.parametrize("model,prior", [(0, "basic"), (1, "alpha"), (2, "alpha_dc"), (6, "eta")])
def test_compact_wire_matches_plain_copy(sim, model, prior):
    """The streamed path ships (steps, choice[, fp32 draw]) records and host threads write the float64 rows
    (ddm_set_host_decode): bit-identical to the kernel's own float64 rows, for every two-column layout, with
    timeouts (max_steps not a multiple of 6), both timeout conventions, float32 rows, ragged chunks and
    odd thread counts."""
    from bayesflow_nddms_b200 import priors

    params = priors.draw_prior_batch(prior, 203, np.random.default_rng(model))
    kw = dict(dt=0.01, max_steps=47, seed=11, dataset_offset=5)
    try:
        for flags in (0, 1, F_F32):
            sim.set_pipeline(1 << 60, -1)
            base = sim.simulate(model, params, 131, flags=flags, **kw)
            assert (base[..., 0] == 0).any() or model in (0, 6)  # timeouts present (signed-rt layouts mark them 0)
            sim.set_pipeline(1, 131 * 17)
            for threads in (-1, 1, 3, 0):
                sim.set_host_decode(threads)
                again = sim.simulate(model, params, 131, flags=flags, **kw)
                assert np.array_equal(base, again), (flags, threads)
                assert sim.last_stats()["n_timeouts"] > 0
    finally:
        sim.set_pipeline(-1, -1)
        sim.set_host_decode(0)


def test_result_arrays_come_from_a_pinned_pool(sim):
    """Mid-size results are allocated in page-locked memory (full-rate copies) and the block is reused once the
    previous result is dropped; a view keeps its block alive; beyond the pool's limit results are plain arrays."""
    import gc

    from bayesflow_nddms_b200 import priors
    from bayesflow_nddms_b200.simulator import _PinnedBlock

    def owner(a):
        while isinstance(a, np.ndarray):
            a = a.base
        return a

    params = priors.draw_prior_batch("basic", 64, np.random.default_rng(1))
    a = sim.simulate(0, params, 500, seed=4, dataset_offset=0)                   # 512 KB
    assert isinstance(owner(a), _PinnedBlock) and a.flags.writeable and a.flags.c_contiguous
    ptr_a = a.ctypes.data
    keep = a[3, :7].copy()
    view = a[3]
    del a
    gc.collect()
    b = sim.simulate(0, params, 500, seed=5, dataset_offset=0)                   # the view still owns the first block
    assert b.ctypes.data != ptr_a and np.array_equal(view[:7], keep)
    ref = b.copy()
    del view, b
    gc.collect()
    c = sim.simulate(0, params, 500, seed=5, dataset_offset=0)
    assert c.ctypes.data in (ptr_a, ref.ctypes.data) or isinstance(owner(c), _PinnedBlock)
    assert np.array_equal(c, ref)
    small = sim.simulate(0, params[:2], 10, seed=5, dataset_offset=0)            # tiny: ordinary array
    assert not isinstance(owner(small), _PinnedBlock)
    # a pool at its limit hands out nothing (the caller then falls back to an ordinary array)
    from bayesflow_nddms_b200.simulator import _PinnedResultPool
    assert _PinnedResultPool(sim._lib, limit=0).empty((64, 1000, 2), np.float64) is None


def test_dc_scaling_is_exact_in_fp32(sim):
    """simulations/Basic_DDM_simulations.py:164-209: (boundary, drift, dc) and (2b, 2d, 2dc) have the
    same choice-RT law; scaling by 2 is exact in binary floating point, so with the same
    Philox stream the trials are identical."""
    a = sim.simulate(0, [1.5, 1.2, 0.5, 0.35, 1.0], 4000, seed=8, dataset_offset=0)
    b = sim.simulate(0, [3.0, 2.4, 0.5, 0.35, 2.0], 4000, seed=8, dataset_offset=0)
    assert np.array_equal(a, b)


# --------------------------------------------------------------------------------------------
# fp64 validation mode on the Philox stream vs the oracle on the same (ideal) stream
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,params,kw", PROD_CASES[:6] + PROD_CASES[8:])
def test_fp64_philox_mode_vs_oracle(sim, oracle, model, params, kw):
    n = 500
    out = sim.simulate(model, params, n, precision=64, flags=F_STEPS, seed=42, dataset_offset=3, **kw)[0]
    steps = sim.last_steps(n)
    o = oracle.simulate_philox(model, params, n, 42, dataset=3, **kw)
    same = steps == o.n_steps
    assert same.mean() >= 0.998          # libdevice vs glibc transcendental ulps can flip a tie
    assert np.array_equal(out[same, 0].view(np.uint64), o.sim_data[same, 0].view(np.uint64))
    assert np.allclose(out[same, 1], o.sim_data[same, 1], rtol=0, atol=1e-12)


# --------------------------------------------------------------------------------------------
# Edge cases
# --------------------------------------------------------------------------------------------
def test_edge_shapes(sim):
    p = np.array([[1.0, 1.2, 0.5, 0.3, 1.0]])
    assert sim.simulate(0, p, 0).shape == (1, 0, 2)
    assert sim.simulate(0, np.empty((0, 5)), 10).shape == (0, 10, 2)
    one = sim.simulate(0, p, 1)
    assert one.shape == (1, 1, 2) and one[0, 0, 0] >= 0.3
    # max_steps = 0: no step is taken, every trial is a timeout at rt = tau
    z = sim.simulate(0, p, 33, max_steps=0)
    assert np.all(z[0, :, 0] == 0.3) and np.all(z[0, :, 1] == 0)
    # max_steps only truncates: the stream of a trial does not depend on it (multiples of the
    # 6-normal Philox block and partial last blocks alike)
    a = sim.simulate(0, p, 300, max_steps=400, seed=5, dataset_offset=0, flags=F_STEPS)
    sa = sim.last_steps(300)
    for ms in (399, 396, 7, 6, 5, 1):
        b = sim.simulate(0, p, 300, max_steps=ms, seed=5, dataset_offset=0, flags=F_STEPS)
        sb = sim.last_steps(300)
        assert sim.last_stats()["used_persistent"] == 1
        assert np.array_equal(np.minimum(sa, ms), sb), ms
        keep = sa < ms
        assert np.array_equal(a[0, keep], b[0, keep]), ms
        assert np.all((b[0, ~keep, 1] == 0) | (sa[~keep] == ms))
    # start point on / outside a boundary: zero steps (beta = 1 -> evidence >= boundary)
    c = sim.simulate(0, [[1.0, 1.2, 1.0, 0.3, 1.0]], 5)
    assert np.all(c[0, :, 0] == 0.3) and np.all(c[0, :, 1] == 1)
    d = sim.simulate(0, [[1.0, 1.2, 0.0, 0.3, 1.0]], 5)
    assert np.all(d[0, :, 0] == 0.3) and np.all(d[0, :, 1] == -1)


def test_zero_and_negative_diffusion_coefficient(sim, oracle):
    """dc == 0: the reference runs a deterministic drift to the boundary; dc < 0 mirrors the noise.  The
    production kernel's state unit does not exist there; the library falls back to the reference formulas."""
    p = np.array([[2.0, 1.0, 0.5, 0.3, 0.0], [1.0, 1.2, 0.4, 0.2, 1.0], [-1.0, 0.8, 0.5, 0.1, 0.0]])
    out = sim.simulate(0, p, 40, seed=3, dataset_offset=0)
    assert sim.last_stats()["used_persistent"] == 0
    o0 = oracle.simulate_philox(0, p[0], 40, 3, dataset=0)
    assert np.all(out[0, :, 1] == 1) and np.allclose(out[0, :, 0], o0.sim_data[:, 0], atol=0.0100001)  # 25 steps of .02, up to a rounding tie
    assert np.all(out[2, :, 1] == -1) and np.all(np.abs(out[2, :, 0] - (0.1 + 40 * 0.01)) < 0.0100001)
    assert set(np.unique(out[1, :, 1])) <= {-1.0, 1.0}
    neg = sim.simulate(0, [[1.0, 1.2, 0.4, 0.2, -1.0]], 2000, seed=3, dataset_offset=1)
    pos = sim.simulate(0, [[1.0, 1.2, 0.4, 0.2, 1.0]], 2000, seed=4, dataset_offset=1)
    assert abs((neg[0, :, 1] > 0).mean() - (pos[0, :, 1] > 0).mean()) < 0.05
    ev = sim.simulate_evidence([[2.0, 1.0, 0.5, 0.3, 0.0, 0.1]], 8, 200, 1, seed=1, dataset_offset=0)
    assert np.all(np.isfinite(ev)) and np.all(ev[0, :, 1] == 1)
    # large batches are scanned for such datasets by several host threads: one dc == 0 anywhere is found
    big = np.tile([1.0, 1.2, 0.4, 0.2, 1.0], (300_000, 1))
    sim.simulate(0, big, 2, seed=3, dataset_offset=0)
    assert sim.last_stats()["used_persistent"] == 1
    for where in (0, 177_777, 299_999):
        big[where, 4] = 0.0
        out = sim.simulate(0, big, 2, seed=3, dataset_offset=0)
        assert sim.last_stats()["used_persistent"] == 0 and np.all(np.isfinite(out))
        big[where, 4] = 1.0


def test_timeout_flag_and_float32_output(sim):
    p = [0.05, 4.0, 0.5, 0.3, 0.3]
    a = sim.simulate(0, p, 256, seed=2, dataset_offset=0)
    b = sim.simulate(0, p, 256, seed=2, dataset_offset=0, flags=F_TIMEOUT1)
    to = a[0, :, 1] == 0
    assert to.sum() > 20
    assert np.all(b[0, to, 1] == 1) and np.array_equal(a[0, ~to], b[0, ~to])
    assert np.all(a[0, to, 0] == 400 * 0.01 + 0.3)
    f = sim.simulate(0, p, 256, seed=2, dataset_offset=0, flags=F_F32)
    assert f.dtype == np.float32 and np.array_equal(f, a.astype(np.float32))


def test_argument_errors(sim):
    with pytest.raises(ValueError):
        sim.simulate(0, np.ones((2, 7)), 10)            # wrong parameter count
    with pytest.raises(ValueError):
        sim.simulate(9, np.ones((2, 5)), 10)            # unknown model
    with pytest.raises(ValueError):
        sim.simulate(0, np.ones((2, 5)), 10, dt=0.0)
    with pytest.raises(ValueError):
        sim.simulate(0, np.ones((2, 5)), 10, precision=16)
    with pytest.raises(ValueError, match="cannot be less than zero"):
        sim.simulate_trialwise([0, 0], [1.0, -0.5], [[3.0, 0.5, 0.4, 1.0]])
    with pytest.raises(ValueError):
        sim.simulate_trialwise([0, 1], [1.0, 0.5], [[3.0, 0.5, 0.4, 1.0]])  # group out of range
    # the context survives errors
    assert sim.simulate(0, np.ones((1, 5)), 4).shape == (1, 4, 2)


def test_trialwise_ragged_groups(sim, oracle):
    from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as stahl

    subj, pe = stahl.synthetic_stahl_like()
    assert subj.size == 19374 and np.unique(subj).size == 89
    data, part_ids, part_index = stahl.impute_dataset(subj, pe, simulator=sim, seed=77)
    assert data.shape == (19374, 2) and data.dtype == np.float64
    alpha_like, alphas = stahl.boundaries_from_pe(pe)
    assert np.array_equal(data[:, 1], alpha_like)
    assert (alphas == 0).sum() > 0
    # bound == 0 -> +ter, zero steps
    pp = stahl.draw_participant_params(89, np.random.default_rng(2024))
    cr = stahl.impute_choicert(part_index, alphas, pp, simulator=sim, seed=77)
    z = alphas == 0
    assert np.array_equal(cr[z], pp[part_index[z], 2])
    # same trials through the oracle on the ideal Philox stream (fp64 mode)
    cr64 = sim.simulate_trialwise(part_index, alphas, pp, seed=77, precision=64, trial_offset=0)[:, 0]
    pick = np.random.default_rng(0).choice(19374, 300, replace=False)
    for i in pick:
        o = oracle.simulate_philox(5, pp[part_index[i]], 1, 77, dataset=0, trial_offset=int(i), bound_in=[alphas[i]])
        assert abs(o.sim_data[0, 0] - cr64[i]) < 1e-12 or abs(abs(o.sim_data[0, 0]) - abs(cr64[i])) <= 0.0100001
    sizes = [b['sim_data'].shape[1] for b in stahl.participant_batches(data, part_index, 89)]
    assert sum(sizes) == 19374 and min(sizes) >= 13


def test_dlpack_handoff_to_torch(sim):
    import torch

    from bayesflow_nddms_b200 import basic_ddm_dc as m

    params = m.batch_draw_prior(64)
    host = m.batch_simulate_trials(params, 500, sim, seed=4, dataset_offset=0)
    dev = m.batch_simulate_trials_device(params, 500, sim, seed=4, dataset_offset=0)
    assert dev.shape == (64, 500, 2) and dev.dtype_bits == 32
    t = torch.from_dlpack(dev)
    assert t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (64, 500, 2)
    assert np.array_equal(t.cpu().numpy(), host.astype(np.float32))
    with pytest.raises(RuntimeError):
        dev.__dlpack__()                       # one-shot
    # the consumer owns the buffer: a second batch must not alias it
    dev2 = m.batch_simulate_trials_device(params, 500, sim, seed=5, dataset_offset=0)
    t2 = torch.from_dlpack(dev2)
    assert t2.data_ptr() != t.data_ptr()
    assert np.array_equal(t.cpu().numpy(), host.astype(np.float32))
    del t, t2
    # configurators: numpy and device-resident agree
    d = m.generative_model(32, sim)
    c = m.configurator(d)
    assert c['summary_conditions'].dtype == np.float32 and c['direct_conditions'].shape == (32, 1)
    dd = m.generative_model(32, sim, device=True)
    cd = m.device_configurator(dd)
    assert cd['summary_conditions'].is_cuda and cd['parameters'].shape == (32, 5)
    assert torch.allclose(cd['direct_conditions'].cpu(), torch.full((32, 1), float(np.log(dd['sim_non_batchable_context']))))


def test_reference_signatures(sim):
    from bayesflow_nddms_b200 import basic_ddm_dc as m0
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1
    from bayesflow_nddms_b200 import set_default_simulator

    set_default_simulator(sim)
    try:
        p = m0.draw_prior()
        assert p.shape == (5,) and p.dtype == np.float64
        out = m0.simulate_trials(p, 300)
        assert out.shape == (300, 2) and out.dtype == np.float64
        assert set(np.unique(out[:, 1])) <= {-1.0, 0.0, 1.0}
        rt, choice = m0.diffusion_trial(3.0, 1.5, 0.5, 0.4, 1.0)
        assert rt >= 0.4 and choice in (-1, 0, 1)
        p1 = m1.draw_prior()
        assert p1.shape == (7,)
        for fn, pp in ((m1.simulate_trials, p1), (m1.simulate_trials_alt, m1.draw_prior_alt()),
                       (m1.simulate_trials_scale, m1.draw_prior_scale()), (m1.simulate_trials_scale2, p1),
                       (m1.simulate_trials_fine, p1)):
            o = fn(pp, 123)
            assert o.shape == (123, 2) and np.all(np.isfinite(o))
        cr, ext = m1.diffusion_trial(3.0, 1.5, 0.5, 0.4, 1.0, 1.0, 0.1)
        assert np.isfinite(cr) and np.isfinite(ext)
        # successive calls advance the dataset counter: fresh streams
        a, b = m0.simulate_trials(p, 50), m0.simulate_trials(p, 50)
        assert not np.array_equal(a, b)
    finally:
        set_default_simulator(None)


def test_device_replay_feed(sim):
    import torch

    from bayesflow_nddms_b200 import basic_ddm_dc as m
    from bayesflow_nddms_b200.replay import DeviceReplayBuffer, replay_iterations

    buf = DeviceReplayBuffer(4)
    n = 0
    for batch in replay_iterations(m, 32, 10, buf, simulator=sim):
        n += 1
        x, y, c = batch['summary_conditions'], batch['parameters'], batch['direct_conditions']
        assert x.is_cuda and y.is_cuda and c.is_cuda and x.dtype == torch.float32
        assert x.shape[0] == 32 and x.shape[2] == 2 and 60 <= x.shape[1] <= 300 and y.shape == (32, 5)
        assert abs(float(c[0, 0]) - np.log(x.shape[1])) < 1e-5 and torch.isfinite(x).all()
    assert n == 10 and len(buf) == 4 and buf.stored_total == 10
    ptrs = {b['summary_conditions'].data_ptr() for b in buf._slots}
    assert len(ptrs) == 4                                # four live device batches, none aliased


def test_reference_wiring_through_bayesflow_simulation_api(sim, monkeypatch):
    """The reference's own wiring (basic_ddm_dc.py:130-134, single_trial_alpha_not_scaled.py:160-164)
    driven through a stand-in of bf.simulation (BayesFlow is not installed here): both the reference's
    non-batched mode and BayesFlow's batched mode, then the reference's configurator."""
    import sys

    import fake_bayesflow
    from bayesflow_nddms_b200 import basic_ddm_dc as m0
    from bayesflow_nddms_b200 import set_default_simulator
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

    bf, bfsim = fake_bayesflow.as_module()
    monkeypatch.setitem(sys.modules, "bayesflow", bf)
    monkeypatch.setitem(sys.modules, "bayesflow.simulation", bfsim)
    set_default_simulator(sim)
    try:
        for m, P in ((m0, 5), (m1, 7)):
            # exactly the reference's lines
            prior = bf.simulation.Prior(prior_fun=m.draw_prior)
            experimental_context = bf.simulation.ContextGenerator(non_batchable_context_fun=m.prior_N)
            simulator = bf.simulation.Simulator(simulator_fun=m.simulate_trials, context_generator=experimental_context)
            generative_model = bf.simulation.GenerativeModel(prior, simulator)
            d = generative_model(8)
            n = d['sim_non_batchable_context']
            assert d['prior_draws'].shape == (8, P) and d['sim_data'].shape == (8, n, 2) and 60 <= n <= 300
            c = m.configurator(d)
            assert c['summary_conditions'].shape == (8, n, 2) and c['parameters'].shape == (8, P)
            # the batched wiring the package offers
            for batched in (True, False):
                gm = m.make_bayesflow_generative_model(batched=batched)
                d = gm(32)
                n = d['sim_non_batchable_context']
                assert d['prior_draws'].shape == (32, P) and d['sim_data'].shape == (32, n, 2)
                assert np.all(np.isfinite(d['sim_data'])) and d['sim_data'].dtype == np.float64
    finally:
        set_default_simulator(None)


@pytest.mark.parametrize("B,N", [(3, 100_003), (100_003, 3), (1, 1_000_000), (7104 * 3 + 1, 33)])
def test_odd_shapes_every_trial_written_once(sim, B, N):
    """Tiles, ragged last tiles and warps that outnumber the work: every trial written exactly once
    (step counts are positive wherever the start point is inside the boundaries), counters consistent."""
    rng = np.random.default_rng(B + N)
    from bayesflow_nddms_b200 import priors

    params = priors.draw_prior_batch("basic", B, rng)
    out = sim.simulate(0, params, N, seed=21, dataset_offset=5, flags=F_STEPS | F_F32)
    steps = sim.last_steps(B * N).reshape(B, N)
    st = sim.last_stats()
    assert st["n_trials"] == B * N and st["total_steps"] == int(steps.sum(dtype=np.int64))
    assert steps.min() >= 1 and steps.max() <= 400
    tau = params[:, 3][:, None]
    assert np.array_equal(out[..., 0], (steps * 0.01 + tau).astype(np.float32))
    assert np.all(np.isin(out[..., 1], (-1.0, 0.0, 1.0))) and st["n_timeouts"] == int((out[..., 1] == 0).sum())
    # a second, differently tiled run of a slice reproduces it
    lo = min(B - 1, 2)
    again = sim.simulate(0, params[lo:lo + 1], N, seed=21, dataset_offset=5 + lo, flags=F_F32)
    assert np.array_equal(again[0], out[lo])


@pytest.mark.parametrize("model,prior", [(0, "basic"), (1, "alpha"), (2, "alpha_dc"), (3, "alpha_scale"), (4, "alpha_scale2"),
                                         (6, "eta")])
def test_fp64_shared_increments_on_random_prior_draws(sim, oracle, model, prior):
    """Breadth: 40 parameter vectors from each model's prior, one launch per model on a shared MT19937
    increment buffer; every dataset must equal the reference loop bit for bit (fp64, both columns, steps)."""
    from bayesflow_nddms_b200 import priors

    rng = np.random.default_rng(1000 + model)
    P = priors.draw_prior_batch(prior, 40, rng)
    n = 64
    kw = dict(dt=0.01, max_steps=400) if model % 2 == 0 else dict(dt=0.001, max_steps=4000)
    normals_all, offs_all, refs, ref_steps = [], [], [], []
    base = 0
    for d in range(P.shape[0]):
        o = oracle.simulate_mt(model, P[d], n, seed=5000 + d, **kw)
        z = oracle.mt_normals(5000 + d, int(o.n_steps.sum()) + 80 * n + 64)
        ob = oracle.simulate_buffer(model, P[d], n, z, **kw)
        assert np.array_equal(ob.sim_data, o.sim_data)
        used = int(ob.consumed.sum())
        off = np.zeros(n, np.int64)
        off[1:] = np.cumsum(ob.consumed)[:-1]
        normals_all.append(z[:used])
        offs_all.append(off + base)
        base += used
        refs.append(o.sim_data)
        ref_steps.append(o.n_steps)
    sim.set_normals_debug(np.concatenate(normals_all), np.concatenate(offs_all))
    try:
        out = sim.simulate(model, P, n, precision=64, flags=F_STEPS, seed=1, dataset_offset=0, **kw)
        steps = sim.last_steps(P.shape[0] * n).reshape(P.shape[0], n)
        st = sim.last_stats()
    finally:
        sim.set_normals_debug(None, None)
    ref = np.stack(refs)
    assert np.array_equal(out.view(np.uint64), ref.view(np.uint64))
    assert np.array_equal(steps, np.stack(ref_steps)) and st["debug_overruns"] == 0


def test_full_size_run_is_reproducible_and_counter_keyed(sim):
    """Size-independent properties at 2e7 trials: two runs agree in every counter and in a checksum of the
    output; a different seed or dataset offset changes it; the f64 and f32 outputs describe the same trials."""
    from bayesflow_nddms_b200 import priors

    B, N = 20_000, 1000
    params = priors.draw_prior_batch("sweep", B, np.random.default_rng(3))

    def run(seed, off):
        out = sim.simulate(0, params, N, seed=seed, dataset_offset=off, dt=1e-3, max_steps=4000, flags=F_F32)
        st = sim.last_stats()
        chk = (float(out[..., 0].sum(dtype=np.float64)), float(out[..., 1].sum(dtype=np.float64)))
        return st, chk

    a, ca = run(5, 0)
    b, cb = run(5, 0)
    for k in ("total_steps", "n_timeouts", "n_upper", "n_trials"):
        assert a[k] == b[k]
    assert ca == cb
    c, cc = run(6, 0)
    d, cd = run(5, B)
    assert cc != ca and cd != ca and c["total_steps"] != a["total_steps"] and d["total_steps"] != a["total_steps"]
    # RT checksum equals dt * total_steps (tau = 0 in the sweep prior) up to float32 row rounding
    assert abs(ca[0] - 1e-3 * a["total_steps"]) < 1e-6 * a["total_steps"]
    assert abs(ca[1] - (2 * a["n_upper"] + a["n_timeouts"] - B * N)) < 0.5


def test_two_contexts_on_two_host_threads(sim):
    """include/ddm_b200.h: one ctx per host thread, functions on distinct ctx are concurrency-safe (ctypes
    releases the GIL around the calls).  Two threads with their own contexts reproduce the serial results."""
    import threading

    import bayesflow_nddms_b200 as pkg
    from bayesflow_nddms_b200 import priors

    params = priors.draw_prior_batch("alpha", 200, np.random.default_rng(8))
    want = [sim.simulate(1, params, 300, seed=100 + i, dataset_offset=7) for i in range(2)]
    got, errs = [None, None], []

    def work(i):
        try:
            with pkg.DDMSimulator(device=0, seed=100 + i) as s:
                for _ in range(5):
                    got[i] = s.simulate(1, params, 300, dataset_offset=7)
        except Exception as e:  # surfaced below
            errs.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errs, errs
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


@pytest.mark.parametrize("model,prior,kw", [(0, "basic", dict(dt=0.001, max_steps=4000)), (0, "basic", dict(dt=0.01, max_steps=400)),
                                            (1, "alpha", dict(dt=0.01, max_steps=400)), (2, "alpha_dc", dict(dt=0.01, max_steps=400)),
                                            (6, "eta", dict(dt=0.001, max_steps=4000))])
def test_check1_at_scale_fp32_kernel_vs_fp64_reference_arithmetic_on_the_same_increments(sim, oracle, model, prior, kw):
    """Check #1 on 1e6 trials per model: the fp32 production kernel against the reference's fp64 loop
    (the validation kernel, itself bit-equal to the CPU oracle) consuming the production kernel's own
    fp32 normals.  Crossing steps and choices must agree except for boundary ties; the tie rate is
    asserted (and printed) here: it is the documented exception of the parity claim."""
    from bayesflow_nddms_b200 import priors

    B, N = 1000, 1000
    P = priors.draw_prior_batch(prior, B, np.random.default_rng(77 + model))
    a = sim.simulate(model, P, N, seed=13, dataset_offset=0, flags=F_STEPS, **kw)
    sa = sim.last_steps(B * N)
    assert sim.last_stats()["used_persistent"] == 1
    b = sim.simulate(model, P, N, seed=13, dataset_offset=0, precision=64, flags=F_STEPS | 32, **kw)
    sb = sim.last_steps(B * N)
    a2, b2 = a.reshape(-1, 2), b.reshape(-1, 2)
    same = sa == sb
    tie_rate = 1.0 - same.mean()
    print(f"model {model} dt={kw['dt']}: {int((~same).sum())} of {B * N} crossing steps differ (tie rate {tie_rate:.2e})")
    assert tie_rate < 5e-5            # measured 1e-6 .. 1.4e-5
    # ... and every one of them is a boundary tie: at the step where the paths part, the reference's fp64
    # evidence (CPU oracle on the exported increments) is within TIE_ULPS fp32 ulps of the boundary
    ulps = []
    for idx in np.flatnonzero(~same)[:64]:
        d, t = divmod(int(idx), N)
        ulps.append(_assert_boundary_tie(sim, oracle, model, P[d], d, t, sa[idx], sb[idx], 13, kw))
    if ulps:
        print(f"    distance of the fp64 evidence from the boundary at the parting step: max {max(ulps):.2f}, "
              f"median {np.median(ulps):.2f} fp32 ulps")
    # where the step count agrees, the reported RT is bit-identical and the choice equal
    assert np.array_equal(a2[same, 0].view(np.uint64), b2[same, 0].view(np.uint64))
    if model in (0, 6):
        assert np.array_equal(a2[same, 1], b2[same, 1])
    else:
        # per-trial redraws decide in fp32 vs fp64: a candidate within rounding of 0 may flip (rarer still)
        ext_same = np.abs(a2[same, 1] - b2[same, 1]) < 1e-4 * (1 + np.abs(b2[same, 1]))
        assert ext_same.mean() > 1 - 1e-4
    # a tie changes a trial's crossing step, not the law: the step-count difference has no drift
    d = (sa.astype(np.int64) - sb.astype(np.int64))[~same]
    if d.size > 20:
        assert abs(np.mean(np.sign(d))) < 0.5


# --------------------------------------------------------------------------------------------
# Check #1 for DDM_MODEL_TRIALWISE (imputation_from_stahl_not_scaled.py:120-148) at its default precision 32
# --------------------------------------------------------------------------------------------
def _stahl_like_bounds(rng, n):
    """Boundaries shaped like the reference's (z(pre_Pe) + 3) / 3 clipped at 0 (imputation...:82-105), with the
    clipped zeros and a band of tiny positive values (a start point within rounding of both boundaries)."""
    b = np.clip((rng.standard_normal(n) * 1.05 + 3.0) / 3.0, 0.0, None)
    b[rng.random(n) < 0.01] = 0.0
    tiny = rng.random(n) < 0.01
    b[tiny] = 10.0 ** rng.uniform(-6, -2, tiny.sum())
    return b


TRIALWISE_CASES = [
    ([3.0, 0.5, 0.4, 1.0], dict(dt=0.01, max_steps=400)),
    ([-1.2, 0.42, 0.25, 0.8], dict(dt=0.001, max_steps=4000)),
    ([0.1, 0.55, 0.3, 0.25], dict(dt=0.01, max_steps=400)),          # small dc: many timeouts
]


@pytest.mark.parametrize("gp,kw", TRIALWISE_CASES)
def test_trialwise_fp32_vs_reference_loop_on_exported_increments(sim, oracle, gp, kw):
    """The product path of the Stahl imputation (precision 32) against the reference's pure-Python loop (CPU
    oracle, fp64) on the kernel's own exported increments: crossing steps and choices equal except boundary
    ties (at most 1 of 384, classified), reported RTs then bit-identical, final states within 1e-5."""
    n, seed, toff = 384, 31, 1000
    rng = np.random.default_rng(int(abs(gp[0]) * 100))
    bounds = _stahl_like_bounds(rng, n)
    bounds[:4] = [0.0, 1e-7, 1e-4, 2.5]
    group = np.zeros(n, np.int32)
    out = sim.simulate_trialwise(group, bounds, [gp], seed=seed, trial_offset=toff, precision=32, flags=F_STEPS, **kw)
    steps = sim.last_steps(n)
    state = sim.simulate_trialwise(group, bounds, [gp], seed=seed, trial_offset=toff, precision=32, flags=F_STATE, **kw)[:, 1]
    assert np.array_equal(out[:, 1], bounds)
    differ = 0
    for t in range(n):
        r, _ = _reference_trial(sim, oracle, 5, gp, 0, toff + t, steps[t], seed, kw, bound_in=[bounds[t]])
        if int(r.n_steps[0]) != steps[t]:
            differ += 1
            _assert_boundary_tie(sim, oracle, 5, gp, 0, toff + t, steps[t], r.n_steps[0], seed, kw, bound_in=[bounds[t]])
            continue
        assert np.sign(out[t, 0]) == r.choice[0], t
        assert out[t, 0].view(np.uint64) == r.sim_data[0, 0].view(np.uint64), t
        # the fp32 state's error scales with the state: a trial with a tiny boundary ends after one step, a whole
        # increment (~ dc sqrt(dt)) beyond it
        assert abs(state[t] - r.evidence[0]) <= 1e-5 * max(bounds[t], abs(r.evidence[0]), 1e-3), (t, bounds[t], state[t], r.evidence[0])
    assert differ <= 1, f"{differ} of {n} crossing steps differ"
    assert np.all(steps[bounds == 0] == 0) and np.all(out[bounds == 0, 0] == gp[2])   # :131-133: zero steps, +ter


def test_trialwise_check1_at_scale(sim, oracle):
    """1e6 Stahl-shaped trials over 89 participants: the precision-32 product path vs the reference's fp64
    arithmetic on the same fp32 normals (DDM_FLAG_F32_NORMALS), incl. bound == 0 and tiny bounds."""
    from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as stahl

    n, G, seed = 1_000_000, 89, 17
    rng = np.random.default_rng(5)
    pp = stahl.draw_participant_params(G, np.random.default_rng(2024))
    group = rng.integers(0, G, n).astype(np.int32)
    bounds = _stahl_like_bounds(rng, n)
    kw = dict(dt=0.01, max_steps=400)
    a = sim.simulate_trialwise(group, bounds, pp, seed=seed, precision=32, flags=F_STEPS, trial_offset=0, **kw)
    sa = sim.last_steps(n)
    b = sim.simulate_trialwise(group, bounds, pp, seed=seed, precision=64, flags=F_STEPS | 32, trial_offset=0, **kw)
    sb = sim.last_steps(n)
    same = sa == sb
    tie_rate = 1.0 - same.mean()
    print(f"trialwise: {int((~same).sum())} of {n} crossing steps differ (tie rate {tie_rate:.2e})")
    assert tie_rate < 5e-5
    assert np.array_equal(a[same, 0].view(np.uint64), b[same, 0].view(np.uint64))
    assert np.array_equal(a[:, 1], bounds) and np.array_equal(b[:, 1], bounds)
    for idx in np.flatnonzero(~same)[:64]:
        i = int(idx)
        _assert_boundary_tie(sim, oracle, 5, pp[group[i]], 0, i, sa[i], sb[i], seed, kw, bound_in=[bounds[i]])


def test_trialwise_calls_consume_fresh_randomness(sim):
    """the reference's per-row loop (imputation...:205-213) draws new noise on every call of
    diffusion_trial; the drop-in must too (a per-simulator trial counter), and stay reproducible on request."""
    from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as stahl
    from bayesflow_nddms_b200 import set_default_simulator

    set_default_simulator(sim)
    try:
        rts = {stahl.diffusion_trial(1.0, 1.4, 0.5, 0.3, 1.0) for _ in range(12)}
        assert len(rts) >= 6                              # 12 identical values before the fix
        part_index = np.zeros(500, np.int32)
        alphas = np.full(500, 1.2)
        pp = np.array([[1.0, 0.5, 0.3, 1.0]])
        x = stahl.impute_choicert(part_index, alphas, pp, simulator=sim)
        y = stahl.impute_choicert(part_index, alphas, pp, simulator=sim)
        assert not np.array_equal(x, y)
        u = stahl.impute_choicert(part_index, alphas, pp, simulator=sim, seed=5, trial_offset=0)
        v = stahl.impute_choicert(part_index, alphas, pp, simulator=sim, seed=5, trial_offset=0)
        assert np.array_equal(u, v)
        # one launch over 500 trials == 500 single-trial calls at the same counters
        w = np.array([sim.simulate_trialwise([0], [1.2], pp, seed=5, trial_offset=i)[0, 0] for i in range(40)])
        assert np.array_equal(w, u[:40])
    finally:
        set_default_simulator(None)


# --------------------------------------------------------------------------------------------
# Round 2: the tile-staged persistent kernel
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", [0, 1, 2, 3, 4, 6])
def test_tile_kernel_equals_round1_persistent_kernel(sim, model):
    """Both schedulers of the same trials (ddm_set_kernel_variant) give the same bits, for host rows, float32
    rows and the compact wire, over tile sizes that do and do not divide the dataset and thresholds that force
    stragglers (a tile recycled while its slow trials still run: they write their own rows)."""
    from bayesflow_nddms_b200 import priors

    name = ["basic", "alpha", "alpha_dc", "alpha_scale", "alpha_scale2", None, "eta"][model]
    params = priors.draw_prior_batch(name, 53, np.random.default_rng(40 + model))
    try:
        for kw in (dict(dt=0.01, max_steps=400), dict(dt=0.001, max_steps=1000)):
            sim.set_kernel_variant(1)
            sim.set_tuning(0, 0, 0)
            ref = sim.simulate(model, params, 307, seed=9, dataset_offset=3, flags=F_STEPS, **kw)
            ref_steps, ref_st = sim.last_steps(53 * 307), sim.last_stats()
            assert ref_st["scheduler"] == 1 and ref_st["tile"] == 64
            sim.set_kernel_variant(0)
            for thr, bps, tile in [(0, 0, 0), (1, 1, 5), (3, 2, 33), (32, 0, 128), (8, 0, 64), (2, 1, 1000)]:
                sim.set_tuning(thr, bps, tile)
                out = sim.simulate(model, params, 307, seed=9, dataset_offset=3, flags=F_STEPS, **kw)
                steps, st = sim.last_steps(53 * 307), sim.last_stats()
                assert st["used_persistent"] == 1 and st["scheduler"] == 2 and st["tile"] <= 128
                assert np.array_equal(out.view(np.uint64), ref.view(np.uint64)), (kw, thr, bps, tile)
                assert np.array_equal(steps, ref_steps)
                for k in ("total_steps", "n_timeouts", "n_upper", "reject_cap_hits", "n_trials"):
                    assert st[k] == ref_st[k], (k, thr, bps, tile)
            sim.set_tuning(2, 1, 16)   # few warps, small tiles: most tiles are recycled with trials still running
            f32 = sim.simulate(model, params, 307, seed=9, dataset_offset=3, flags=F_F32, **kw)
            assert np.array_equal(f32, ref.astype(np.float32))
            # the latency kernel (small launches' default: one thread per trial, speculative six-step blocks)
            for variant in (2, -1):
                sim.set_kernel_variant(variant)
                sim.set_tuning(0, 0, 0)
                out = sim.simulate(model, params, 307, seed=9, dataset_offset=3, flags=F_STEPS, **kw)
                steps, st = sim.last_steps(53 * 307), sim.last_stats()
                assert st["used_persistent"] == 1 and st["scheduler"] == 3, variant
                assert np.array_equal(out.view(np.uint64), ref.view(np.uint64)), (kw, variant)
                assert np.array_equal(steps, ref_steps)
                for k in ("total_steps", "n_timeouts", "n_upper", "reject_cap_hits", "n_trials"):
                    assert st[k] == ref_st[k], (k, variant)
                f32 = sim.simulate(model, params, 307, seed=9, dataset_offset=3, flags=F_F32, **kw)
                assert np.array_equal(f32, ref.astype(np.float32))
    finally:
        sim.set_kernel_variant(-1)
        sim.set_tuning(0, 0, 0)


def test_latency_kernel_edges_and_routing(sim):
    """The latency kernel against the tile kernel where the block structure shows: max_steps that is not a multiple of
    six (and of twelve: two blocks per iteration), 0, 1 and 5; trials that start on a boundary; the trialwise (Stahl)
    model; and the automatic choice -- launches of at most 256 Ki trials take it, larger ones the tile kernel."""
    from bayesflow_nddms_b200 import priors
    from bayesflow_nddms_b200.imputation_from_stahl_not_scaled import synthetic_stahl_like, boundaries_from_pe, draw_participant_params

    P = priors.draw_prior_batch("sweep", 40, np.random.default_rng(77))
    P[0, 2] = 0.0            # starts on the lower boundary: no step
    P[1, 2] = 1.0            # ... on the upper one
    P[2, 1] = 1e-3           # a tiny boundary: crosses at the first step
    try:
        for ms in (0, 1, 5, 6, 7, 11, 12, 13, 400, 401, 407):
            res = {}
            for variant in (0, 2):
                sim.set_kernel_variant(variant)
                out = sim.simulate(0, P, 211, dt=0.01, max_steps=ms, seed=3, dataset_offset=5, flags=F_STEPS)
                res[variant] = (out.copy(), sim.last_steps(40 * 211).copy(), sim.last_stats())
            assert res[0][2]["scheduler"] == 2 and res[2][2]["scheduler"] == 3
            assert np.array_equal(res[0][0].view(np.uint64), res[2][0].view(np.uint64)), ms
            assert np.array_equal(res[0][1], res[2][1]) and res[2][1].max() <= ms
            for k in ("total_steps", "n_timeouts", "n_upper"):
                assert res[0][2][k] == res[2][2][k], (k, ms)
        # trialwise model
        subj, pe = synthetic_stahl_like(np.random.default_rng(3), nsubs=12, ntrials_total=2500)
        _, alphas = boundaries_from_pe(pe)
        alphas[:3] = 0.0
        _, idx = np.unique(subj, return_inverse=True)
        pp = draw_participant_params(12, np.random.default_rng(4))
        tw = {}
        for variant in (0, 2, -1):
            sim.set_kernel_variant(variant)
            tw[variant] = sim.simulate_trialwise(idx, alphas, pp, trial_offset=123).copy()
            assert sim.last_stats()["scheduler"] == (2 if variant == 0 else 3)
        assert np.array_equal(tw[0].view(np.uint64), tw[2].view(np.uint64)) and np.array_equal(tw[0].view(np.uint64), tw[-1].view(np.uint64))
        # automatic choice by size
        sim.set_kernel_variant(-1)
        sim.run(0, P, 6553, 0.01, 400)                       # 262 120 trials <= 256 Ki
        assert sim.last_stats()["scheduler"] == 3
        sim.run(0, P, 6554, 0.01, 400)
        assert sim.last_stats()["scheduler"] == 2
        for bad in (3, -2):
            with pytest.raises(ValueError):
                sim.set_kernel_variant(bad)
    finally:
        sim.set_kernel_variant(-1)


def test_tile_kernel_general_model_and_wire(sim):
    """Three-column rows (DDM_MODEL_GENERAL, both output styles) and the compact wire through the tile kernel."""
    from bayesflow_nddms_b200 import priors, two_channel

    rng = np.random.default_rng(3)
    raw = np.column_stack([rng.normal(0, 2, 31), 1.0 + rng.random(31), np.full(31, 0.5), np.full(31, 0.3), 0.5 + rng.random(31),
                           0.6 + rng.random(31), np.full(31, 0.3), np.full(31, 0.5), np.full(31, 0.6), np.full(31, 0.2), np.full(31, 0.1)])
    for P in (two_channel.canonical_drift_dc5(raw), two_channel.canonical_alpha_dc(raw)):   # both output styles' cousins
        try:
            sim.set_kernel_variant(1)
            sim.set_tuning(0, 0, 0)
            ref = sim.simulate(7, P, 211, seed=4, dataset_offset=0)
            sim.set_kernel_variant(0)
            for thr, bps, tile in [(0, 0, 0), (2, 1, 16), (8, 0, 96)]:
                sim.set_tuning(thr, bps, tile)
                assert np.array_equal(sim.simulate(7, P, 211, seed=4, dataset_offset=0).view(np.uint64), ref.view(np.uint64))
        finally:
            sim.set_kernel_variant(-1)
            sim.set_tuning(0, 0, 0)
    params = priors.draw_prior_batch("alpha", 97, np.random.default_rng(6))
    try:
        sim.set_pipeline(1 << 60, -1)
        base = sim.simulate(1, params, 131, seed=2, dataset_offset=0, dt=0.01, max_steps=61)
        sim.set_pipeline(1, 131 * 10)
        sim.set_tuning(2, 1, 8)
        again = sim.simulate(1, params, 131, seed=2, dataset_offset=0, dt=0.01, max_steps=61)
        assert np.array_equal(base, again)
    finally:
        sim.set_pipeline(-1, -1)
        sim.set_tuning(0, 0, 0)


def test_trialwise_persistent_equals_generic_bitwise(sim):
    """DDM_MODEL_TRIALWISE at precision 32 now runs on the persistent tile kernel (VERDICT r1 missing #4):
    same bits as the one-thread-per-trial kernel, ragged tails, zero and tiny boundaries included."""
    from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as stahl

    rng = np.random.default_rng(12)
    n, G = 100_003, 89
    pp = stahl.draw_participant_params(G, np.random.default_rng(2024))
    group = rng.integers(0, G, n).astype(np.int32)
    bounds = _stahl_like_bounds(rng, n)
    for kw in (dict(dt=0.01, max_steps=400), dict(dt=0.001, max_steps=777)):
        sim.set_kernel_variant(0)
        try:
            a = sim.simulate_trialwise(group, bounds, pp, seed=3, trial_offset=11, flags=F_STEPS, **kw)
        finally:
            sim.set_kernel_variant(-1)
        sa, st = sim.last_steps(n), sim.last_stats()
        assert st["used_persistent"] == 1 and st["scheduler"] == 2
        c = sim.simulate_trialwise(group, bounds, pp, seed=3, trial_offset=11, flags=F_STEPS, **kw)   # default at this size: latency kernel
        assert sim.last_stats()["scheduler"] == 3 and np.array_equal(a.view(np.uint64), c.view(np.uint64))
        assert np.array_equal(sa, sim.last_steps(n))
        b = sim.simulate_trialwise(group, bounds, pp, seed=3, trial_offset=11, flags=F_STEPS | F_GENERIC, **kw)
        sb, st2 = sim.last_steps(n), sim.last_stats()
        assert st2["used_persistent"] == 0
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)) and np.array_equal(sa, sb)
        for k in ("total_steps", "n_timeouts", "n_upper", "n_trials"):
            assert st[k] == st2[k]
        f32 = sim.simulate_trialwise(group, bounds, pp, seed=3, trial_offset=11, flags=F_F32, **kw)
        assert np.array_equal(f32, a.astype(np.float32))
    # a participant without noise (dc == 0) has no state unit: the library keeps the reference's formulas
    pp0 = pp.copy()
    pp0[5, 3] = 0.0
    out = sim.simulate_trialwise(group[:5000], bounds[:5000], pp0, seed=3, trial_offset=0)
    assert sim.last_stats()["used_persistent"] == 0 and np.all(np.isfinite(out))


def test_global_dataset_index_is_64_bit(sim, oracle):
    """dataset_counter grows without bound; the index is now 64-bit (low word = Philox
    counter word 3, the rest in the stream word = word 0), a launch must not straddle a multiple of 2^32 and the
    Python counter skips to the next multiple instead."""
    p = np.array([[1.0, 1.2, 0.5, 0.3, 1.0]] * 3)
    base = sim.simulate(0, p, 50, seed=5, dataset_offset=7)
    far = sim.simulate(0, p, 50, seed=5, dataset_offset=(9 << 32) + 7)
    again = sim.simulate(0, p, 50, seed=5, dataset_offset=(9 << 32) + 7)
    assert not np.array_equal(base, far) and np.array_equal(far, again)
    big = (9 << 32) + 8
    z = sim.export_normals(big, 4, 0, 0, 60, seed=5, precision=64)
    assert np.max(np.abs(z - oracle.philox_normals(5, big, 4, 0, 0, 60))) < 1e-13
    assert not np.array_equal(z, sim.export_normals(8, 4, 0, 0, 60, seed=5, precision=64))
    with pytest.raises(ValueError, match="straddle"):
        sim.simulate(0, p, 50, seed=5, dataset_offset=(1 << 32) - 2)
    with pytest.raises(ValueError):
        sim.simulate(0, p, 50, seed=5, dataset_offset=1 << 56)
    keep = sim.dataset_counter
    try:
        sim.dataset_counter = (1 << 32) - 2
        out = sim.simulate(0, p, 50, seed=5)                       # rolls to 2^32 instead of failing
        assert sim.dataset_counter == (1 << 32) + 3
        assert np.array_equal(out, sim.simulate(0, p, 50, seed=5, dataset_offset=1 << 32))
    finally:
        sim.dataset_counter = keep


def test_simulate_histogram_is_the_histogram_of_the_rows(sim):
    """ddm_simulate_histogram (C5 as SURVEY 8d specifies it: host parameters in, device-reduced histogram out) returns
    exactly the histogram of the rows the same call leaves resident, for both row layouts, and agrees with the kernel's
    counters; the rows remain available for DLPack hand-off."""
    import torch

    from bayesflow_nddms_b200 import basic_ddm_dc as m0
    from bayesflow_nddms_b200 import priors
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

    for mod, prior, basic in ((m0, "sweep", True), (m1, "alpha", False)):
        P = priors.draw_prior_batch(prior, 300, np.random.default_rng(5))
        kw = dict(dt=1e-3, max_steps=4000) if basic else dict(dt=0.01, max_steps=400)
        h = mod.batch_simulate_histogram(P, 700, sim, seed=4, dataset_offset=9, n_bins=97, rt_max=3.3, **kw)
        st = sim.last_stats()
        rows = torch.from_dlpack(sim.last_output_dlpack()).cpu().numpy().astype(np.float64).reshape(-1, 2)
        assert rows.shape[0] == 300 * 700
        if basic:
            rt, choice = rows[:, 0], np.sign(rows[:, 1])
        else:
            rt, choice = np.abs(rows[:, 0]), np.sign(rows[:, 0])
        b = rt * (97 / 3.3)
        inside = b < 97
        for sign, key in ((1, "upper"), (-1, "lower")):
            want = np.bincount(b[(choice == sign) & inside].astype(np.int64), minlength=97)
            assert np.array_equal(h[key].astype(np.int64), want), key
        assert h["missing"] == int((choice == 0).sum()) == st["n_timeouts"]
        assert h["overflow"] == int(((choice != 0) & ~inside).sum())
        assert int(h["upper"].sum() + h["lower"].sum()) + h["missing"] + h["overflow"] == 300 * 700
        # same trials as the row-returning call
        again = mod.batch_simulate_trials(P, 700, sim, seed=4, dataset_offset=9, flags=F_F32, **kw)
        assert np.array_equal(again.reshape(-1, 2).astype(np.float64), rows)
    with pytest.raises(ValueError):
        sim.simulate_histogram(7, np.zeros((2, 24)), 10)


@pytest.mark.parametrize("prior,basic", [("sweep", True), ("alpha", False)])
def test_streamed_histogram_equals_one_launch(sim, prior, basic):
    """From 64 Mi trials on ddm_simulate_histogram produces the batch in chunks of datasets (parameters of chunk i+1
    uploaded, and rows of chunk i-1 reduced, beside chunk i's kernel).  Forced at a small size with small chunks, it
    returns the histogram, the counters and the resident rows of the single-launch path, bit for bit."""
    import torch

    from bayesflow_nddms_b200 import basic_ddm_dc as m0
    from bayesflow_nddms_b200 import priors
    from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

    mod = m0 if basic else m1
    P = priors.draw_prior_batch(prior, 3000, np.random.default_rng(11))
    kw = dict(dt=1e-3, max_steps=4000) if basic else dict(dt=0.01, max_steps=400)
    try:
        sim.set_pipeline(1 << 62, -1)                       # chunking off
        h1 = mod.batch_simulate_histogram(P, 500, sim, seed=8, dataset_offset=77, n_bins=120, rt_max=3.0, **kw)
        st1 = sim.last_stats()
        rows1 = torch.from_dlpack(sim.last_output_dlpack()).cpu().numpy().copy()
        sim.set_pipeline(0, 40_000)                         # forced, chunks of >= 80 datasets: 188, 1406, 703, ...
        h2 = mod.batch_simulate_histogram(P, 500, sim, seed=8, dataset_offset=77, n_bins=120, rt_max=3.0, **kw)
        st2 = sim.last_stats()
        rows2 = torch.from_dlpack(sim.last_output_dlpack()).cpu().numpy().copy()
    finally:
        sim.set_pipeline(-1, -1)
    assert st2["kernel_launches"] > st1["kernel_launches"] + 4     # it did run in chunks
    for key in ("upper", "lower"):
        assert np.array_equal(h1[key], h2[key]), key
    assert h1["missing"] == h2["missing"] and h1["overflow"] == h2["overflow"]
    for key in ("total_steps", "n_timeouts", "n_upper", "n_trials"):
        assert st1[key] == st2[key], key
    assert rows1.shape == rows2.shape and np.array_equal(rows1, rows2)


@pytest.mark.parametrize("model,prior", [(0, "basic"), (1, "alpha")])
def test_default_transfer_plan_at_a_million_trials(sim, model, prior):
    """Round 2 lowered the streaming thresholds to 1e6 trials (compact records for (rt, choice) rows, two plain chunks
    into a page-locked destination for the layouts with an external column): the defaults return the bits of one launch +
    one copy, into the pool's pinned result array and into a caller's pageable one."""
    from bayesflow_nddms_b200 import priors

    P = priors.draw_prior_batch(prior, 1003, np.random.default_rng(2))
    for flags in (0, F_F32):
        try:
            sim.set_pipeline(1 << 60, -1)
            base = sim.simulate(model, P, 1000, seed=3, dataset_offset=1, flags=flags)
        finally:
            sim.set_pipeline(-1, -1)
        got = sim.simulate(model, P, 1000, seed=3, dataset_offset=1, flags=flags)             # pinned pool array
        st = sim.last_stats()
        assert np.array_equal(base, got)
        assert st["d2h_bytes"] > 0                                                           # it was streamed
        if model == 0:
            assert st["host_decode_threads"] > 0 and st["d2h_bytes"] == 1003 * 1000 * 4          # compact records
        else:
            assert st["host_decode_threads"] == 0                                                # plain chunks at this size
        mine = np.empty_like(base)                                                            # pageable destination
        out = sim.simulate(model, P, 1000, seed=3, dataset_offset=1, flags=flags, out=mine)
        assert out is mine and np.array_equal(base, mine)
