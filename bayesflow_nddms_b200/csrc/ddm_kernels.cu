// ddm_kernels.cu -- hand-written sm_100a kernels of the DDM trial simulator.
//
//   prep_kernel        raw fp64 parameters -> per-dataset fp32 constants
//   persistent_kernel  production path: one lane per trial, persistent warps that
//                      refill finished lanes from a global work counter
//   generic_kernel     one thread per trial, naive scheduling: fp64 validation mode,
//                      shared-increment (debug buffer) mode, the trialwise (Stahl)
//                      variant, and the bit-exact cross-check of the persistent kernel
//   export / philox    parity hooks
//
// Replaces (reference, all Python/numba): basic_ddm_dc.py:85-125,
// single_trial_alpha_not_scaled.py:107-155, :926-974, :1237-1285, :1471-1519,
// :1710-1722, imputation_from_stahl_not_scaled.py:120-148, :205-213.
#include "ddm_kernels.cuh"

namespace ddm {

// --------------------------------------------------------------------------------
// prep: one thread per dataset
// --------------------------------------------------------------------------------
// The per-dataset (per-participant, model 5) constants of the production kernels from one row of parameters: fp64
// arithmetic, rounded once.  prep_kernel runs it once per dataset; the latency kernel, whose launches are too small to
// be worth a kernel of their own in front of them, runs it per trial (same code, same bits).
__device__ __forceinline__ DsConst prep_constants(const double *__restrict__ p, int model, double dt) {
    DsConst c;
#pragma unroll
    for (int i = 0; i < 8; i++) c.v[i] = 0.f;
    const double unit1 = sqrt(dt) * SQRT_2LN2_D;  // noise scale of one step at dc = 1, in lg2 units
    if (model == 0) {  // [drift, boundary, beta, tau, dc]
        const double U = unit1 * p[4];
        c.v[0] = (float)(p[0] * dt / U);
        c.v[1] = (float)(p[1] * (p[2] - 0.5) / U);
        c.v[2] = (float)(0.5 * p[1] / U);
        c.v[3] = (float)U;
    } else if (model == 6) {  // eta: [mu_drift, alpha, beta, ter, eta, dc]
        const double U = unit1 * p[5];
        c.v[0] = (float)(p[0] * dt / U);
        c.v[1] = (float)(p[1] * (p[2] - 0.5) / U);
        c.v[2] = (float)(0.5 * p[1] / U);
        c.v[3] = (float)U;
        c.v[4] = (float)(p[4] * dt / U);
    } else if (model == 5) {  // trialwise groups: [drift, beta, ter, dc] in the BOUND layout (boundary supplied per trial)
        const double U = unit1 * p[3];
        c.v[0] = (float)(p[0] * dt / U);
        c.v[1] = (float)((p[1] - 0.5) / U);
        c.v[4] = (float)(0.5 / U);
        c.v[7] = (float)U;
    } else if (model == 2) {  // alt: [drift, alpha, beta, ter, std_dc, mu_dc, sigma1]
        c.v[0] = (float)(p[0] * dt / unit1);
        c.v[1] = (float)(p[1] * (p[2] - 0.5) / unit1);
        c.v[2] = (float)(0.5 * p[1] / unit1);
        c.v[3] = (float)p[5];
        c.v[4] = (float)p[4];
        c.v[5] = (float)p[6];
        c.v[7] = (float)unit1;
    } else {  // alpha / scale / scale2: [drift, mu_alpha, beta, ter, std_alpha, dc, sigma1(, gamma)]
        const double U = unit1 * p[5];
        c.v[0] = (float)(p[0] * dt / U);
        c.v[1] = (float)((p[2] - 0.5) / U);
        c.v[2] = (float)p[1];
        c.v[3] = (float)p[4];
        c.v[4] = (float)(0.5 / U);
        c.v[5] = (float)p[6];
        c.v[6] = (model == 3) ? (float)p[7] : (model == 4 ? 2.f : 1.f);
        c.v[7] = (float)U;
    }
    return c;
}

__global__ void prep_kernel(const double *__restrict__ params, DsConst *__restrict__ dconst,
                            uint32_t n_datasets, uint32_t n_params, int model, double dt) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_datasets) return;
    dconst[d] = prep_constants(params + (size_t)d * n_params, model, dt);
}

__global__ void prep_general_kernel(const double *__restrict__ params, GenConst *__restrict__ gconst, uint32_t n_datasets,
                                    double dt) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_datasets) return;
    const double *p = params + (size_t)d * 24;
    const double unit1 = sqrt(dt) * SQRT_2LN2_D;
    GenConst c;
#pragma unroll
    for (int i = 0; i < 24; i++) c.v[i] = 0.f;
    c.v[0] = (float)(p[0] * dt / unit1);
    c.v[1] = (float)(p[1] * dt / unit1);
    c.v[2] = (float)p[2];
    c.v[3] = (float)p[3];
    c.v[4] = (float)p[4];
    c.v[5] = (float)p[5];
    c.v[6] = (float)((p[6] - 0.5) / unit1);
    c.v[7] = (float)(0.5 / unit1);
    c.v[8] = (float)p[0];
    c.v[9] = (float)p[1];
    c.v[10] = (float)unit1;
    const int n_ext = (int)p[21];
    for (int ch = 0; ch < 2; ch++) {
        const double *e = p + 8 + 6 * ch;
        float *o = c.v + 11 + 5 * ch;
        if (ch < n_ext) {
            o[0] = (float)(-e[4] / e[5]);
            o[1] = (float)(e[0] / e[5]);
            o[2] = (float)(e[1] / e[5]);
            o[3] = (float)(e[2] / e[5]);
            o[4] = (float)(e[3] / e[5]);
        }
    }
    c.v[21] = (float)p[22];
    gconst[d] = c;
}

// --------------------------------------------------------------------------------
// output store
// --------------------------------------------------------------------------------
template <bool OUT64>
__device__ __forceinline__ void store_triple(void *out, uint64_t idx, double o0, double o1, double o2) {
    if (OUT64) {
        double *o = reinterpret_cast<double *>(out) + 3 * idx;
        o[0] = o0; o[1] = o1; o[2] = o2;
    } else {
        float *o = reinterpret_cast<float *>(out) + 3 * idx;
        o[0] = (float)o0; o[1] = (float)o1; o[2] = (float)o2;
    }
}

template <bool OUT64>
__device__ __forceinline__ void store_pair(void *out, uint64_t idx, double o0, double o1) {
    if (OUT64) {
        reinterpret_cast<double2 *>(out)[idx] = make_double2(o0, o1);
    } else {
        reinterpret_cast<float2 *>(out)[idx] = make_float2((float)o0, (float)o1);
    }
}

// --------------------------------------------------------------------------------
// persistent kernel
// --------------------------------------------------------------------------------
// Work item = (dataset, tile of <= `tile` consecutive trials).  A warp claims items
// with one atomicAdd each and hands the tile's trials to lanes that need one.  All
// lanes then advance in lock-step, six Euler steps per Philox block; lanes whose
// trial has crossed are frozen by predication.  When at least `refill_threshold`
// lanes are frozen the warp takes the (divergent, so deliberately batched) finish +
// refill path.  Philox counters are (step block, trial, dataset): results do not
// depend on which lane / warp / SM / GPU ran a trial.
#ifndef DDM_PERSISTENT_BLOCK
#define DDM_PERSISTENT_BLOCK 256  // threads per block (A/B: 128 and 512 measured equal within 0.5 %)
#endif
#ifndef DDM_PERSISTENT_MIN_BLOCKS
#define DDM_PERSISTENT_MIN_BLOCKS (1536 / DDM_PERSISTENT_BLOCK)
#endif
template <int KIND, bool OUT64>
__global__ void __launch_bounds__(DDM_PERSISTENT_BLOCK, DDM_PERSISTENT_MIN_BLOCKS) persistent_kernel(const RunArgs a) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;

    // warp-uniform tile cursor
    uint32_t cur = 0, end = 0, tile_ds = 0;
    bool more = true;
    DsConst tile_c;
#pragma unroll
    for (int i = 0; i < 8; i++) tile_c.v[i] = 0.f;

    // per-lane trial
    TrialF32 t;
    t.x = 0.f; t.h = 0.f; t.c0 = 0.f; t.u = 0.f; t.ext = 0.f; t.ext2 = 0.f;
    float x = 0.f;
    uint32_t n = 0, blk = 0, trial = 0, ds = 0;
    uint32_t p = 0;    // 1 = stepping
    bool has = false;  // holds a trial (stepping, or finished and waiting to be emitted)

    unsigned long long acc_steps = 0;
    uint32_t acc_timeouts = 0, acc_upper = 0, acc_cap = 0;

    const int thr = a.refill_threshold;
    constexpr bool BASIC = (KIND == KIND_FIXED || KIND == KIND_DRIFT);

    for (;;) {
        // ---- finish: emit every frozen trial ---------------------------------------
        if (has && p == 0u) {
            int choice = (x >= t.h) ? 1 : ((x <= -t.h) ? -1 : 0);
            // Lanes always run whole 6-step blocks.  A trial that was still inside the boundaries
            // after max_steps steps is a timeout whatever it did in the surplus steps of its last
            // block (max_steps need not be a multiple of 6).
            if (n > a.max_steps) { n = a.max_steps; choice = 0; }
            const double tau = (KIND == KIND_GENERAL) ? 0.0 : a.params[(size_t)ds * a.n_params + 3];
            double o0, o1;
            const uint64_t idx = (uint64_t)ds * a.n_trials + trial;
            if (!DDM_CHECK(a.stats, ds < a.n_datasets && trial < a.n_trials && n <= a.max_steps)) { has = false; continue; }
            if (KIND == KIND_GENERAL) {
                const double tau_g = a.params[(size_t)ds * a.n_params + 7];
                const bool style0 = a.gconst[ds].v[21] == 0.f;
                if (style0) {
                    trial_outputs<true>(a.flags, choice, n, a.dt, tau_g, 0.0, o0, o1);
                    store_triple<OUT64>(a.out, idx, o0, o1, (double)t.ext);
                } else {
                    trial_outputs<false>(a.flags, choice, n, a.dt, tau_g, (double)t.ext, o0, o1);
                    store_triple<OUT64>(a.out, idx, o0, o1, (double)t.ext2);
                }
            } else {
                if (a.flags & FLAG_WIRE_COMPACT) {  // chunked host pipeline: the host writes the rows (ddm_wire.cu)
                    if (BASIC) reinterpret_cast<int32_t *>(a.out)[idx] = wire_pack(n, choice);
                    else reinterpret_cast<int2 *>(a.out)[idx] = make_int2(wire_pack(n, choice), __float_as_int(t.ext));
                } else {
                    trial_outputs<BASIC>(a.flags, choice, n, a.dt, tau, (double)t.ext, o0, o1);
                    store_pair<OUT64>(a.out, idx, o0, o1);
                }
            }
            if (a.steps_out) a.steps_out[idx] = (int32_t)n;
            acc_steps += n;
            acc_timeouts += (choice == 0);
            acc_upper += (choice > 0);
            has = false;
        }
        // ---- refill: hand out trials of the current tile, claiming tiles as needed ---
        for (;;) {
            const unsigned empty = __ballot_sync(FULL_MASK, !has);
            if (empty == 0u) break;
            if (cur == end) {
                if (!more) break;
                unsigned long long w = 0;
                if (lane == 0) w = atomicAdd(a.work_counter, 1ull);
                w = __shfl_sync(FULL_MASK, w, 0);
                if (w >= a.n_items) { more = false; break; }
                uint32_t ti = 0;
                if (a.tiles_per_dataset == 1u) {  // many datasets: a claim is a whole dataset, no division
                    tile_ds = (uint32_t)w;
                } else {
                    tile_ds = (uint32_t)w / a.tiles_per_dataset;  // host keeps n_items < 2^32
                    ti = (uint32_t)w - tile_ds * a.tiles_per_dataset;
                }
                cur = ti * a.tile;
                end = min(cur + a.tile, a.n_trials);
                const float4 *src = reinterpret_cast<const float4 *>((KIND == KIND_GENERAL) ? (const void *)(a.gconst + tile_ds)
                                                                                            : (const void *)(a.dconst + tile_ds));
                const float4 c0 = __ldg(src), c1 = __ldg(src + 1);
                tile_c.v[0] = c0.x; tile_c.v[1] = c0.y; tile_c.v[2] = c0.z; tile_c.v[3] = c0.w;
                tile_c.v[4] = c1.x; tile_c.v[5] = c1.y; tile_c.v[6] = c1.z; tile_c.v[7] = c1.w;
            }
            const uint32_t rank = __popc(empty & lt_mask);
            const uint32_t avail = end - cur;
            if (!has && rank < avail) {
                ds = tile_ds;
                trial = cur + rank;
                if (KIND == KIND_GENERAL)
                    trial_setup_general(a.gconst[ds], trial + a.trial_offset, ds + a.dataset_offset, a.key, t, acc_cap);
                else
                    trial_setup_f32<KIND>(tile_c, trial + a.trial_offset, ds + a.dataset_offset, a.key, t, acc_cap);
                x = t.x;
                n = 0;
                blk = 0;
                has = true;
                p = ((fabsf(x) < t.h) && (a.max_steps > 0u)) ? 1u : 0u;
            }
            cur += min((uint32_t)__popc(empty), avail);
        }
        if (!__any_sync(FULL_MASK, has)) break;
        // once the work has run out there is nothing to refill with: run the warp's last trials to the end
        const int thr_now = (more || cur != end) ? thr : 32;

        // ---- step: tight, branch-free inner loop (round keys and constants stay in uniform registers)
        unsigned alive = __ballot_sync(FULL_MASK, p != 0u);
        const int live_min = 32 - thr_now;  // keep stepping while more than this many lanes are alive
        do {
            Normals6Scaled z;
            philox_pairs_lg2(blk, trial + a.trial_offset, ds + a.dataset_offset, STREAM_STEP, a.key, z);
            euler6_warp(x, n, alive, t.c0, t.h, z, a.max_steps);
            blk++;
        } while (__popc(alive) > live_min);
        p = (alive >> lane) & 1u;
    }

    // ---- per-warp statistics --------------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_steps += __shfl_xor_sync(FULL_MASK, acc_steps, o);
        acc_timeouts += __shfl_xor_sync(FULL_MASK, acc_timeouts, o);
        acc_upper += __shfl_xor_sync(FULL_MASK, acc_upper, o);
        acc_cap += __shfl_xor_sync(FULL_MASK, acc_cap, o);
    }
    if (lane == 0) {
        atomicAdd(a.stats + STAT_STEPS, acc_steps);
        atomicAdd(a.stats + STAT_TIMEOUTS, (unsigned long long)acc_timeouts);
        atomicAdd(a.stats + STAT_UPPER, (unsigned long long)acc_upper);
        if (acc_cap) atomicAdd(a.stats + STAT_REJECT_CAP, (unsigned long long)acc_cap);
    }
}


// --------------------------------------------------------------------------------
// tile kernel (production)
// --------------------------------------------------------------------------------
// The persistent kernel above pays for every finished trial inside the divergent finish + refill pass: the
// per-trial set-up (an aux Philox block, three Box-Muller pairs, the redraw loop) and the fp64 output arithmetic
// run with only the refilled lanes active.  At the reference's own dt = .01 a trial lasts ~7 blocks and that
// pass was more than half of all instructions (round 1: 39 lane-instructions per Euler step against 17 in the
// stepping loop).  Here both ends of a trial are done a whole tile at a time with all 32 lanes busy:
//   claim    a warp claims a tile of <= tile_cap consecutive trials of one dataset and sets all of them up at
//            once (lane i takes trials i, i + 32, ...), staging each trial's start state (x0, h, c0) and its
//            external columns in shared memory (north_star: "per-dataset parameters staged in shared memory");
//   finish   a finished lane parks (steps << 2 | choice + 1) in the tile's result slot: one STS;
//   refill   ... and takes the tile's next trial: a ballot/popc prefix and two or three LDS -- right inside the stepping
//            loop (the "fast path", ~45 issue slots, no loop re-entry) as long as the tile can serve every finished lane;
//   flush    when the next tile but one is claimed the warp turns the 4-byte results into output rows --
//            fp64 arithmetic of the reference, lane-contiguous vector stores.
// A warp works on two tiles at a time (the one being handed out and the previous one, draining).  When a third
// is claimed while the oldest still has trials running (first-passage times are heavy-tailed), that tile is
// flushed as far as it has got and its stragglers write their own rows when they finish ("direct").
// Philox counters, set-up and step arithmetic are those of the generic kernel: results are bit-identical.
constexpr uint32_t SLOT_EMPTY = 0xffffffffu;   // the lane holds no trial
constexpr uint32_t SLOT_DIRECT = 0xfffffffeu;  // its tile's buffer was recycled: the lane emits its own row
constexpr uint32_t CODE_PENDING = 0xffffffffu; // result slot of a trial that has not finished (choice + 1 is never 3)

template <int KIND>
struct TileLayout {
    static constexpr bool XH = KIND == KIND_BOUND || KIND == KIND_DC || KIND == KIND_GENERAL || KIND == KIND_TRIALWISE;
    static constexpr bool C0 = KIND == KIND_DC || KIND == KIND_GENERAL || KIND == KIND_DRIFT || KIND == KIND_TRIALWISE;
    static constexpr bool EXT = KIND == KIND_BOUND || KIND == KIND_DC || KIND == KIND_GENERAL;
    static constexpr bool EXT2 = KIND == KIND_GENERAL;
    // 32-bit words per tile slot: result codes and external columns are double-buffered, the start state is not
    static constexpr uint32_t WORDS = 2u + (XH ? 2u : 0u) + (C0 ? 1u : 0u) + (EXT ? 2u : 0u) + (EXT2 ? 2u : 0u);
    // slots per tile buffer, a compile-time constant so that every buffer is the warp's base address plus an
    // immediate: the largest tile that still leaves room for six resident blocks of eight warps per SM
    static constexpr uint32_t T = WORDS <= 7u ? 128u : 96u;
};

template <bool OUT64>
__device__ __forceinline__ void store_col(void *out, uint64_t at, double v) {
    if (OUT64) reinterpret_cast<double *>(out)[at] = v;
    else reinterpret_cast<float *>(out)[at] = (float)v;
}

struct TileStats {
    unsigned long long steps;
    uint32_t timeouts, upper;
};

// Per-warp bookkeeping that only the tile-change path needs lives in shared memory next to the tile buffers
// (registers are for the stepping loop): the draining tile's identity and the warp's statistics.
enum TileMeta : uint32_t {
    M_ODS = 0, M_OFIRST, M_OCOUNT,            // the tile before the current one: dataset, first trial, trials
    M_STEPS = 4,                              // 64-bit, two words
    M_TIMEOUTS = 6, M_UPPER, M_CAP,
    M_WORDS = 16
};

// One finished trial -> its output row.  COLS: 0 = the whole row; 1 = only the columns that depend on the
// trial's outcome (a straggler whose external columns the partial flush has already written).
template <int KIND, bool OUT64, int COLS>
__device__ __forceinline__ void tile_emit(const RunArgs &a, uint64_t idx, uint32_t c, uint32_t ds, float ext, float ext2,
                                          TileStats &st) {
    constexpr bool BASIC = (KIND == KIND_FIXED || KIND == KIND_DRIFT);
    const uint32_t n = c >> 2;
    const int choice = (int)(c & 3u) - 1;
    if (!DDM_CHECK(a.stats, idx < (uint64_t)a.n_datasets * a.n_trials && ds < a.n_datasets && n <= a.max_steps && (c & 3u) != 3u)) return;
    double o0, o1;
    if (KIND == KIND_GENERAL) {
        const double tau = a.params[(size_t)ds * a.n_params + 7];
        if (a.gconst[ds].v[21] == 0.f) {  // (rt, choice, ext1)
            trial_outputs<true>(a.flags, choice, n, a.dt, tau, 0.0, o0, o1);
            store_col<OUT64>(a.out, 3 * idx, o0);
            store_col<OUT64>(a.out, 3 * idx + 1, o1);
            if (COLS == 0) store_col<OUT64>(a.out, 3 * idx + 2, (double)ext);
        } else {                          // (signed rt, ext1, ext2)
            trial_outputs<false>(a.flags, choice, n, a.dt, tau, 0.0, o0, o1);
            store_col<OUT64>(a.out, 3 * idx, o0);
            if (COLS == 0) {
                store_col<OUT64>(a.out, 3 * idx + 1, (double)ext);
                store_col<OUT64>(a.out, 3 * idx + 2, (double)ext2);
            }
        }
    } else if (a.flags & FLAG_WIRE_COMPACT) {  // chunked host pipeline: the host writes the rows (ddm_wire.cpp)
        if (BASIC) reinterpret_cast<int32_t *>(a.out)[idx] = (int32_t)c;
        else if (COLS == 0) reinterpret_cast<int2 *>(a.out)[idx] = make_int2((int32_t)c, __float_as_int(ext));
        else reinterpret_cast<int32_t *>(a.out)[2 * idx] = (int32_t)c;
    } else {
        double tau, e = (double)ext;
        if (KIND == KIND_TRIALWISE) {
            tau = a.params[(size_t)a.group[idx] * 4 + 2];
            e = a.bound_in[idx];
        } else {
            tau = a.params[(size_t)ds * a.n_params + 3];
        }
        trial_outputs<BASIC>(a.flags, choice, n, a.dt, tau, e, o0, o1);
        if (BASIC || COLS == 0) store_pair<OUT64>(a.out, idx, o0, o1);
        else store_col<OUT64>(a.out, 2 * idx, o0);
    }
    if (a.steps_out) a.steps_out[idx] = (int32_t)n;
    st.steps += n;
    st.timeouts += (choice == 0);
    st.upper += (choice > 0);
}

// The external columns of a trial that is still running when its tile's buffer is recycled.
template <int KIND, bool OUT64>
__device__ __forceinline__ void tile_emit_ext_only(const RunArgs &a, uint64_t idx, uint32_t ds, float ext, float ext2) {
    if (KIND == KIND_GENERAL) {
        if (a.gconst[ds].v[21] == 0.f) {
            store_col<OUT64>(a.out, 3 * idx + 2, (double)ext);
        } else {
            store_col<OUT64>(a.out, 3 * idx + 1, (double)ext);
            store_col<OUT64>(a.out, 3 * idx + 2, (double)ext2);
        }
    } else if (KIND == KIND_TRIALWISE) {
        store_col<OUT64>(a.out, 2 * idx + 1, a.bound_in[idx]);
    } else if (KIND == KIND_BOUND || KIND == KIND_DC) {
        if (a.flags & FLAG_WIRE_COMPACT) reinterpret_cast<int32_t *>(a.out)[2 * idx + 1] = __float_as_int(ext);
        else store_col<OUT64>(a.out, 2 * idx + 1, (double)ext);
    }
}

// All 32 lanes: the rows of the tile described by meta[0..2] (dataset, first trial, trials) from its result
// slots, statistics into the warp's counters.  partial: trials whose slot is still CODE_PENDING get their
// external columns only (they will write the rest themselves when they finish).
template <int KIND, bool OUT64>
__device__ __forceinline__ void tile_flush(const RunArgs &a, const uint32_t *code, const float *ext, const float *ext2,
                                        uint32_t *stats, uint32_t ds, uint32_t first, uint32_t count, bool partial) {
    using L = TileLayout<KIND>;
    TileStats st{0ull, 0u, 0u};
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t idx0 = (uint64_t)ds * a.n_trials + first;
    if (!DDM_CHECK(a.stats, count <= TileLayout<KIND>::T && (uint64_t)first + count <= a.n_trials)) return;
    for (uint32_t i = lane; i < count; i += 32u) {
        const uint32_t c = code[i];
        const float e1 = L::EXT ? ext[i] : 0.f, e2 = L::EXT2 ? ext2[i] : 0.f;
        if (c != CODE_PENDING) tile_emit<KIND, OUT64, 0>(a, idx0 + i, c, ds, e1, e2, st);
        else if (partial) tile_emit_ext_only<KIND, OUT64>(a, idx0 + i, ds, e1, e2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        st.steps += __shfl_xor_sync(FULL_MASK, st.steps, o);
        st.timeouts += __shfl_xor_sync(FULL_MASK, st.timeouts, o);
        st.upper += __shfl_xor_sync(FULL_MASK, st.upper, o);
    }
    if (lane == 0) {
        *reinterpret_cast<unsigned long long *>(stats) += st.steps;
        stats[M_TIMEOUTS - M_STEPS] += st.timeouts;
        stats[M_UPPER - M_STEPS] += st.upper;
    }
    __syncwarp();
}

// Resident blocks per SM: five blocks of 256 threads at 48 registers for every kind.  At 40 registers (six blocks) ptxas
// either spilled a loop-carried register of the stepping loop (the kinds whose tile set-up draws normals) or rebuilt the
// lane mask inside it (the others); the bare stepping loop is bound by the FMA-heavy pipe and loses 1 % from 12 to 10
// warps per scheduler (profiles/r02_microbench_occupancy.txt), and the A/B has five blocks 1-3 % ahead for every kind.
#ifndef DDM_TILE_MIN_BLOCKS
#define DDM_TILE_MIN_BLOCKS (1280 / DDM_PERSISTENT_BLOCK)
#endif
#ifndef DDM_TILE_MIN_BLOCKS_FIXED
#define DDM_TILE_MIN_BLOCKS_FIXED (1280 / DDM_PERSISTENT_BLOCK)
#endif
template <int KIND>
constexpr int tile_min_blocks() {
    return (KIND == KIND_FIXED || KIND == KIND_DRIFT) ? DDM_TILE_MIN_BLOCKS_FIXED : DDM_TILE_MIN_BLOCKS;
}

// A straggler's own row (its tile's buffer has been recycled).  Not inlined: the fp64 output arithmetic and the
// address computations stay out of the register allocation of the pass around the stepping loop.
template <int KIND, bool OUT64>
__device__ __noinline__ void tile_emit_direct(const RunArgs &a, uint32_t ds, uint32_t trial, uint32_t c, uint32_t *meta) {
    TileStats st{0ull, 0u, 0u};
    tile_emit<KIND, OUT64, 1>(a, (uint64_t)ds * a.n_trials + trial, c, ds, 0.f, 0.f, st);
    atomicAdd(reinterpret_cast<unsigned long long *>(meta + M_STEPS), st.steps);
    if (st.timeouts) atomicAdd(meta + M_TIMEOUTS, st.timeouts);
    if (st.upper) atomicAdd(meta + M_UPPER, st.upper);
}

template <int KIND, bool OUT64>
__global__ void __launch_bounds__(DDM_PERSISTENT_BLOCK, tile_min_blocks<KIND>()) tile_kernel(const __grid_constant__ RunArgs a) {
    using L = TileLayout<KIND>;
    extern __shared__ uint32_t tile_smem[];
    const unsigned lane = threadIdx.x & 31u;
    // The lane's masks and the shared address of the warp's result slots are read in every refill; left alone, ptxas
    // rebuilds them each time from the thread index (S2R, shifts, a multiply-add chain for the shared-memory base)
    // rather than keep three registers.  Values that come out of a volatile asm cannot be rematerialised.
    unsigned lane_bit, lt_mask;
    asm volatile("mov.u32 %0, %%lanemask_eq;" : "=r"(lane_bit));
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    const RngConsts rk = pinned_rng_consts();
    constexpr uint32_t T = L::T;
    uint32_t *meta = tile_smem + (threadIdx.x >> 5) * (L::WORDS * T + M_WORDS);
    uint32_t *code = meta + M_WORDS;                                      // [2][T] result slots
    uint32_t code_s = (uint32_t)__cvta_generic_to_shared(code);           // the same, as a 32-bit shared-window address
    asm volatile("" : "+r"(code_s));
    float *sext = reinterpret_cast<float *>(code + 2u * T);               // [2][T] external column (EXT)
    float *sext2 = sext + (L::EXT ? 2u * T : 0u);                         // [2][T] second external column (EXT2)
    float *sx = sext2 + (L::EXT2 ? 2u * T : 0u);                          // [T] staged start state of the current tile
    float *sh = sx + (L::XH ? T : 0u);
    float *sc0 = sh + (L::XH ? T : 0u);
    if (lane < M_WORDS) meta[lane] = 0u;
    __syncwarp();

    // warp-uniform: the current tile (dataset, first trial, trials, trials handed out, buffer) and lane masks
    uint32_t c_ds = 0, c_first = 0, cn = 0, ci = 0, cb = 0;
    bool more = true;
    unsigned held = 0u;    // lanes that hold a trial (stepping, or finished and not yet parked)
    unsigned direct = 0u;  // ... whose tile buffer has been recycled: they write their own rows
    unsigned alive = 0u;   // ... that are still stepping
    DsConst tile_c;
#pragma unroll
    for (int i = 0; i < 8; i++) tile_c.v[i] = 0.f;

    // per-lane trial
    float x = 0.f, h = 0.f, c0 = 0.f;
    uint32_t n = 0, blk = 0, trial = 0, ds = 0;
    uint32_t slot = 0;  // buffer * T + index within the tile

    const int thr = a.refill_threshold;

    for (;;) {
        // ==== tile change (once per tile): claim, write out the tile before the exhausted one, set the new one up ====
        // (kept outside the loop below: its fp64 output arithmetic and address computations must not compete for
        // registers with the stepping loop, where the ten Philox round keys live in uniform registers)
        if (__any_sync(FULL_MASK, more && ci == cn)) {
            unsigned long long w = 0;
            if (lane == 0) w = atomicAdd(a.work_counter, 1ull);
            w = __shfl_sync(FULL_MASK, w, 0);
            if (w >= a.n_items) {
                more = false;
            } else {
                const uint32_t o_count = meta[M_OCOUNT];
                if (o_count != 0u) {
                    // trials of that tile still running (first-passage times are heavy-tailed) lose their slot: their
                    // external columns are written now, the rest by themselves when they finish
                    const bool mine = (held & ~direct & lane_bit) && ((slot >= T) != (cb != 0u));
                    const unsigned stragglers = __ballot_sync(FULL_MASK, mine);
                    const uint32_t ob = (cb ^ 1u) * T;
                    tile_flush<KIND, OUT64>(a, code + ob, sext + ob, sext2 + ob, meta + M_STEPS, meta[M_ODS], meta[M_OFIRST], o_count,
                                            stragglers != 0u);
                    direct |= stragglers;
                }
                if (lane == 0) { meta[M_ODS] = c_ds; meta[M_OFIRST] = c_first; meta[M_OCOUNT] = cn; }
                cb ^= 1u;
                uint32_t ti = 0;
                if (a.tiles_per_dataset == 1u) {  // a claim is a whole dataset, no division
                    c_ds = (uint32_t)w;
                } else {
                    c_ds = (uint32_t)w / a.tiles_per_dataset;  // host keeps n_items < 2^32
                    ti = (uint32_t)w - c_ds * a.tiles_per_dataset;
                }
                c_first = ti * a.tile;
                cn = min(a.tile, a.n_trials - c_first);
                ci = 0;
                if (KIND != KIND_GENERAL && KIND != KIND_TRIALWISE) {
                    const float4 *src = reinterpret_cast<const float4 *>(a.dconst + c_ds);
                    const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
                    tile_c.v[0] = v0.x; tile_c.v[1] = v0.y; tile_c.v[2] = v0.z; tile_c.v[3] = v0.w;
                    tile_c.v[4] = v1.x; tile_c.v[5] = v1.y; tile_c.v[6] = v1.z; tile_c.v[7] = v1.w;
                }
                // set the whole tile up, all lanes busy
                const uint32_t nb = cb * T;
                uint32_t cap_hits = 0;
                if (!DDM_CHECK(a.stats, cn >= 1u && cn <= T && c_ds < a.n_datasets && c_first + cn <= a.n_trials)) cn = 0;
                for (uint32_t i = lane; i < cn; i += 32u) {
                    if (KIND != KIND_FIXED) {
                        TrialF32 t;
                        t.x = 0.f; t.h = 0.f; t.c0 = 0.f; t.u = 0.f; t.ext = 0.f; t.ext2 = 0.f;
                        if (KIND == KIND_GENERAL) {
                            trial_setup_general(a.gconst[c_ds], c_first + i + a.trial_offset, c_ds + a.dataset_offset, a.key, t, cap_hits);
                        } else if (KIND == KIND_TRIALWISE) {
                            const uint64_t g = (uint64_t)c_first + i;
                            trial_setup_trialwise(a.dconst[a.group[g]], (float)a.bound_in[g], t);
                        } else {
                            trial_setup_f32<KIND>(tile_c, c_first + i + a.trial_offset, c_ds + a.dataset_offset, a.key, t, cap_hits);
                        }
                        if (L::XH) { sx[i] = t.x; sh[i] = t.h; }
                        if (L::C0) sc0[i] = t.c0;
                        if (L::EXT) sext[nb + i] = t.ext;
                        if (L::EXT2) sext2[nb + i] = t.ext2;
                    }
                    code[nb + i] = CODE_PENDING;
                }
                if (cap_hits) atomicAdd(meta + M_CAP, cap_hits);
            }
            __syncwarp();
        }
        // ==== the tile's life: finish / refill / step until it is used up ====
        bool done = false;
        for (;;) {
            // ---- finish: a frozen lane parks its result in its tile's slot (one STS) --------------------------
            // (alive comes out of the stepping loop's inline PTX: the vote below only tells the compiler that it is
            // warp-uniform, which keeps the loop condition -- and with it the whole loop -- provably convergent)
            alive = __ballot_sync(FULL_MASK, (alive & lane_bit) != 0u);
            const unsigned fin = held & ~alive;
            if (fin & lane_bit) {
                int choice = (x >= h) ? 1 : ((x <= -h) ? -1 : 0);
                // lanes always run whole 6-step blocks; a trial still inside the boundaries after max_steps
                // steps is a timeout whatever it did in the surplus steps of its last block
                if (n > a.max_steps) { n = a.max_steps; choice = 0; }
                const uint32_t c = (uint32_t)wire_pack(n, choice);
                if (direct & lane_bit) {  // rare: the trial outlived its tile's buffer
                    tile_emit_direct<KIND, OUT64>(a, ds, trial, c, meta);
                } else if (DDM_CHECK(a.stats, slot < 2u * T)) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(code_s + 4u * slot), "r"(c) : "memory");
                }
            }
            held &= ~fin;
            direct &= ~fin;
            // ---- refill: lanes without a trial take the tile's next ones (two or three LDS) -------------------
            const unsigned need = ~held;
            const uint32_t i = ci + __popc(need & lt_mask);
            const bool take = (need & lane_bit) && i < cn;
            if (take) {
                DDM_CHECK(a.stats, i < T);
                slot = cb * T + i;
                ds = c_ds;
                trial = c_first + i;
                x = L::XH ? sx[i] : tile_c.v[1];
                h = L::XH ? sh[i] : tile_c.v[2];
                c0 = L::C0 ? sc0[i] : tile_c.v[0];
                n = 0;
                blk = 0;
            }
            const unsigned took = __ballot_sync(FULL_MASK, take);
            ci += __popc(took);
            held |= took;
            // lanes that were stepping go on; a fresh trial steps unless it starts on or outside a boundary
            const bool stepping = take ? ((fabsf(x) < h) && (a.max_steps > 0u)) : ((alive & lane_bit) != 0u);
            alive = __ballot_sync(FULL_MASK, stepping);
            // (warp-uniform by construction; taken through a vote so that the compiler knows it, too, and keeps the
            // stepping loop's round keys in uniform registers)
            if (__any_sync(FULL_MASK, more && ci == cn && held != FULL_MASK)) break;  // lanes are waiting, the tile is used up: next tile
            if (__all_sync(FULL_MASK, held == 0u)) { done = true; break; }            // no trial left in the warp and none to be had

            // ---- step, and refill in place -------------------------------------------------------------------
            // The stepping loop (round keys and constants in uniform registers) runs until `thr` of the warp's trials have
            // finished -- all of them once the work has run out.  If the current tile can serve every finished lane and
            // none of them is a straggler, the lanes park their results and take their next trials right here (the fast
            // path: ~30 issue slots, no loop re-entry); only a used-up tile, a straggler's own row or the end of the work
            // leave the loop for the pass above.  Every value that decides control flow in here goes through a vote or a
            // warp reduction first: they are warp-uniform by construction, but derived from a shuffled work index, and
            // with bounds it cannot *prove* uniform ptxas guards the loop's votes with BRA.DIV and reloads the round keys
            // on every block.
            const unsigned held_u = __ballot_sync(FULL_MASK, (held & lane_bit) != 0u);
            const unsigned direct_u = __ballot_sync(FULL_MASK, (direct & lane_bit) != 0u);
            const int live_min = __any_sync(FULL_MASK, more || ci != cn) ? max(__popc(held_u) - thr, 0) : 0;
            uint32_t ci_u = __reduce_max_sync(FULL_MASK, ci);
            const uint32_t cn_u = __reduce_max_sync(FULL_MASK, cn);
            const uint32_t slot0 = cb * T;
            for (;;) {
                do {  // the tight loop: one Philox block, three Box-Muller pairs, six predicated steps
                    Normals6Scaled z;
                    philox_pairs_lg2(blk, trial + a.trial_offset, ds + a.dataset_offset, STREAM_STEP, a.key, rk, z);
                    euler6_warp_lb(x, n, alive, lane_bit, c0, h, z, a.max_steps);
                    blk++;
                } while (__popc(alive) > live_min);  // keep stepping while more lanes than this are alive
                const unsigned fin = held_u & ~alive;
                const uint32_t nf = __popc(fin);
                if (ci_u + nf > cn_u || (fin & direct_u) != 0u) break;
                const bool mine = (fin & lane_bit) != 0u;
                if (mine) {
                    int choice = (x >= h) ? 1 : ((x <= -h) ? -1 : 0);
                    if (n > a.max_steps) choice = 0;  // whole blocks: see the pass
                    if (DDM_CHECK(a.stats, slot < 2u * T))
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(code_s + 4u * slot), "r"((uint32_t)wire_pack(min(n, a.max_steps), choice)) : "memory");
                    const uint32_t i = ci_u + __popc(fin & lt_mask);
                    DDM_CHECK(a.stats, i < cn_u && cn_u <= T);
                    slot = slot0 + i;
                    ds = c_ds;
                    trial = c_first + i;
                    x = L::XH ? sx[i] : tile_c.v[1];
                    h = L::XH ? sh[i] : tile_c.v[2];
                    c0 = L::C0 ? sc0[i] : tile_c.v[0];
                    n = 0;
                    blk = 0;
                }
                ci_u += nf;
                const bool stepping = mine ? ((fabsf(x) < h) && (a.max_steps > 0u)) : ((alive & lane_bit) != 0u);
                alive = __ballot_sync(FULL_MASK, stepping);
            }
            ci = ci_u;
        }
        if (__any_sync(FULL_MASK, done)) break;
    }

    // ---- the last two tiles, then the warp's statistics -------------------------------------------------
    __syncwarp();
    {
        const uint32_t o_count = meta[M_OCOUNT];
        if (o_count != 0u) {
            const uint32_t ob = (cb ^ 1u) * T;
            tile_flush<KIND, OUT64>(a, code + ob, sext + ob, sext2 + ob, meta + M_STEPS, meta[M_ODS], meta[M_OFIRST], o_count, false);
        }
        if (cn != 0u) {
            const uint32_t nb = cb * T;
            tile_flush<KIND, OUT64>(a, code + nb, sext + nb, sext2 + nb, meta + M_STEPS, c_ds, c_first, cn, false);
        }
    }
    if (lane == 0) {
        atomicAdd(a.stats + STAT_STEPS, *reinterpret_cast<unsigned long long *>(meta + M_STEPS));
        atomicAdd(a.stats + STAT_TIMEOUTS, (unsigned long long)meta[M_TIMEOUTS]);
        atomicAdd(a.stats + STAT_UPPER, (unsigned long long)meta[M_UPPER]);
        if (meta[M_CAP]) atomicAdd(a.stats + STAT_REJECT_CAP, (unsigned long long)meta[M_CAP]);
    }
}

// --------------------------------------------------------------------------------
// generic kernel: one thread per trial
// --------------------------------------------------------------------------------
struct NormalStreamBuf {
    const double *z, *end;
    bool overrun;
    __device__ __forceinline__ double next() {
        if (z >= end) { overrun = true; return 0.0; }
        return *z++;
    }
};

template <typename Real, int KIND, bool BUFFER, bool OUT64>
__global__ void __launch_bounds__(128) generic_kernel(const RunArgs a, uint64_t total) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = g < total;
    unsigned long long acc_steps = 0;
    uint32_t tout = 0, upper = 0, cap = 0;

    if (active) {
        uint32_t ds, trial;
        if (KIND == KIND_TRIALWISE) {
            ds = 0;
            trial = (uint32_t)g;
        } else {
            ds = (uint32_t)(g / a.n_trials);
            trial = (uint32_t)(g - (uint64_t)ds * a.n_trials);
        }
        const uint32_t ds_g = ds + a.dataset_offset, trial_g = trial + a.trial_offset;
        const double *prm = (KIND == KIND_TRIALWISE) ? a.params + (size_t)a.group[g] * 4
                                                     : a.params + (size_t)ds * a.n_params;
        const double tau = (KIND == KIND_TRIALWISE) ? prm[2] : prm[3];
        uint32_t n = 0;
        int choice = 0;
        double ext = 0.0, final_ev = 0.0;

        if (sizeof(Real) == 4 && !BUFFER && !(a.flags & FLAG_REFERENCE_ARITHMETIC)) {
            // ---- fp32 / Philox: the persistent kernel's arithmetic, naive scheduling ----
            TrialF32 t;
            if (KIND == KIND_TRIALWISE) {
                trial_setup_trialwise(a.dconst[a.group[g]], (float)a.bound_in[g], t);
            } else {
                const DsConst dc = a.dconst[ds];
                trial_setup_f32<(KIND == KIND_TRIALWISE ? KIND_FIXED : KIND)>(dc, trial_g, ds_g, a.key, t, cap);
            }
            float x = t.x;
            uint32_t p = ((fabsf(x) < t.h) && (a.max_steps > 0u)) ? 1u : 0u;
            for (uint32_t blk = 0; p != 0u; blk++)
                step_block_f32<true>(blk, trial_g, ds_g, a.key, t, x, n, p, a.max_steps);
            choice = (x >= t.h) ? 1 : ((x <= -t.h) ? -1 : 0);
            ext = (KIND == KIND_TRIALWISE) ? a.bound_in[g] : (double)t.ext;
            final_ev = (double)__fmul_rn(__fadd_rn(x, t.h), t.u);
        } else {
            // ---- reference arithmetic in Real (fp64: the reference's exact operation order;
            //      fp32: the same formulas rounded to float), normals from Philox or a buffer ----
            NormalStreamBuf buf{BUFFER ? a.dbg_z + a.dbg_off[g] : nullptr, BUFFER ? a.dbg_z + a.dbg_n : nullptr, false};
            double zc[6];
            uint32_t zblk = 0xffffffffu;
            auto philox_normal = [&](uint32_t stream, uint32_t idx) -> double {
                const uint32_t b = idx / 6u;
                const uint32_t tag = b | (stream << 31);  // cache tag
                if (tag != zblk) {
                    if (sizeof(Real) == 8 && !(a.flags & 32)) {
                        philox_normals6_f64(b, trial_g, ds_g, stream, a.key, zc);
                    } else {
                        float zf[6];
                        philox_normals6_f32(b, trial_g, ds_g, stream, a.key, zf);
#pragma unroll
                        for (int i = 0; i < 6; i++) zc[i] = zf[i];
                    }
                    zblk = tag;
                }
                return zc[idx - 6u * b];
            };
            Real drift, beta, bound, dcoef, sigma1 = 0, gain = 1, latent = 0;
            if (KIND == KIND_TRIALWISE) {
                drift = (Real)prm[0]; beta = (Real)prm[1]; dcoef = (Real)prm[3];
                bound = (Real)a.bound_in[g];
                latent = bound;
            } else if (KIND == KIND_FIXED) {
                drift = (Real)prm[0]; bound = (Real)prm[1]; beta = (Real)prm[2]; dcoef = (Real)prm[4];
            } else if (KIND == KIND_DRIFT) {  // basic_ddm_eta_dc.py:87-88
                bound = (Real)prm[1]; beta = (Real)prm[2]; dcoef = (Real)prm[5];
                const Real z = (Real)(BUFFER ? buf.next() : philox_normal(STREAM_AUX, 1u));
                drift = (Real)prm[0] + (Real)prm[4] * z;
            } else {
                drift = (Real)prm[0]; beta = (Real)prm[2]; sigma1 = (Real)prm[6];
                const Real mu = (KIND == KIND_BOUND) ? (Real)prm[1] : (Real)prm[5];
                const Real sd = (Real)prm[4];
                uint32_t i = 0;
                for (;;) {
                    const Real z = (Real)(BUFFER ? buf.next() : philox_normal(STREAM_AUX, 1u + i));
                    latent = mu + sd * z;  // separate mul and add: see -fmad=false in build
                    i++;
                    if (latent > (Real)0) break;
                    if (i >= 6u * REJECT_CAP_BLOCKS - 1u) { cap++; latent = (Real)1e-30; break; }
                }
                if (KIND == KIND_BOUND) {
                    bound = latent; dcoef = (Real)prm[5];
                    gain = (a.model == 3) ? (Real)prm[7] : ((a.model == 4) ? (Real)2 : (Real)1);
                } else {
                    bound = (Real)prm[1]; dcoef = latent; gain = (Real)1;
                }
            }
            const Real dt = (Real)a.dt, sqrt_dt = (Real)a.sqrt_dt;
            Real ev = bound * beta;
            while ((ev > (Real)0) && (ev < bound) && (n < a.max_steps)) {
                const Real z = (Real)(BUFFER ? buf.next() : philox_normal(STREAM_STEP, n));
                const Real t1 = drift * dt;
                const Real t2 = sqrt_dt * dcoef;
                const Real t3 = t2 * z;
                ev = ev + (t1 + t3);
                n++;
            }
            choice = (ev >= bound) ? 1 : ((ev <= (Real)0) ? -1 : 0);
            final_ev = (double)ev;
            if (KIND == KIND_TRIALWISE) {
                ext = (double)bound;
            } else if (KIND != KIND_FIXED && KIND != KIND_DRIFT) {
                const Real z = (Real)(BUFFER ? buf.next() : philox_normal(STREAM_AUX, 0u));
                const Real e = gain * latent + sigma1 * z;
                ext = (double)e;
            }
            if (BUFFER && buf.overrun) atomicAdd(a.stats + STAT_DBG_OVERRUN, 1ull);
        }
        double o0, o1;
        trial_outputs<(KIND == KIND_FIXED || KIND == KIND_DRIFT)>(a.flags, choice, n, a.dt, tau, ext, o0, o1);
        if (a.flags & 16) o1 = final_ev;
        store_pair<OUT64>(a.out, g, o0, o1);
        if (a.steps_out) a.steps_out[g] = (int32_t)n;
        acc_steps = n;
        tout = (choice == 0);
        upper = (choice > 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_steps += __shfl_xor_sync(FULL_MASK, acc_steps, o);
        tout += __shfl_xor_sync(FULL_MASK, tout, o);
        upper += __shfl_xor_sync(FULL_MASK, upper, o);
        cap += __shfl_xor_sync(FULL_MASK, cap, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(a.stats + STAT_STEPS, acc_steps);
        if (tout) atomicAdd(a.stats + STAT_TIMEOUTS, (unsigned long long)tout);
        if (upper) atomicAdd(a.stats + STAT_UPPER, (unsigned long long)upper);
        if (cap) atomicAdd(a.stats + STAT_REJECT_CAP, (unsigned long long)cap);
    }
}

// --------------------------------------------------------------------------------
// latency kernel: small launches (one trial per lane or fewer -- the reference's own batch sizes: 1 x 300, 64 x 500,
// 19 374 Stahl trials).  With nothing to refill, a launch lasts as long as its longest trial's chain of dependent
// instructions, so this kernel is written for the chain and not for the issue slots: one thread per trial, and a
// six-step block is SPECULATIVE -- the six states x1..x6 are a bare chain of six FADDs, the six boundary tests are
// independent of one another, and the step at which the trial stops (and its state there) is selected afterwards --
// instead of the production loop's predicated chain  FADD -> FSETP -> @p FADD -> ...  whose every step waits for the
// previous step's predicate.  The next block's Philox rounds and Box-Muller pairs do not depend on the state at all and
// run under the current block's steps (two blocks per iteration).  Same Philox counters, same set-up, same fp32
// operations on the states that are kept: results are bit-identical to the tile kernel's (tests).  Lanes run whole
// blocks and the count is clamped at the end, exactly like the tile kernel.
// --------------------------------------------------------------------------------
__device__ __forceinline__ void spec_block(float &x, uint32_t &n, bool &alive, float c0, float h, const Normals6Scaled &z) {
    const float x1 = __fadd_rn(x, __fmaf_rn(z.s[0], z.c[0], c0));
    const float x2 = __fadd_rn(x1, __fmaf_rn(z.s[0], z.sn[0], c0));
    const float x3 = __fadd_rn(x2, __fmaf_rn(z.s[1], z.c[1], c0));
    const float x4 = __fadd_rn(x3, __fmaf_rn(z.s[1], z.sn[1], c0));
    const float x5 = __fadd_rn(x4, __fmaf_rn(z.s[2], z.c[2], c0));
    const float x6 = __fadd_rn(x5, __fmaf_rn(z.s[2], z.sn[2], c0));
    const bool i1 = fabsf(x1) < h, i2 = fabsf(x2) < h, i3 = fabsf(x3) < h, i4 = fabsf(x4) < h, i5 = fabsf(x5) < h,
               i6 = fabsf(x6) < h;
    // steps taken = 1 + the number of leading "still inside" tests among the first five; the state is the one after the last step taken
    const bool a2 = i1 && i2, a3 = a2 && i3, a4 = a3 && i4, a5 = a4 && i5;
    const uint32_t k = 1u + (i1 ? 1u : 0u) + (a2 ? 1u : 0u) + (a3 ? 1u : 0u) + (a4 ? 1u : 0u) + (a5 ? 1u : 0u);
    const float xs = a5 ? x6 : (a4 ? x5 : (a3 ? x4 : (a2 ? x3 : (i1 ? x2 : x1))));
    if (alive) {
        x = xs;
        n += k;
        alive = a5 && i6;
    }
}

#ifndef DDM_LATENCY_BLOCKS
#define DDM_LATENCY_BLOCKS 2  // blocks per iteration
#endif
#ifndef DDM_LATENCY_THREADS
#define DDM_LATENCY_THREADS 128  // threads per block
#endif
template <int KIND, bool OUT64>
__global__ void __launch_bounds__(DDM_LATENCY_THREADS) latency_kernel(const RunArgs a, uint64_t total) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc_steps = 0;
    uint32_t tout = 0, upper = 0, cap = 0;
    if (g < total) {
        uint32_t ds, trial;
        if (KIND == KIND_TRIALWISE) {
            ds = 0;
            trial = (uint32_t)g;
        } else {
            ds = (uint32_t)(g / a.n_trials);
            trial = (uint32_t)(g - (uint64_t)ds * a.n_trials);
        }
        const uint32_t ds_g = ds + a.dataset_offset, trial_g = trial + a.trial_offset;
        const double *prm = (KIND == KIND_TRIALWISE) ? a.params + (size_t)a.group[g] * 4 : a.params + (size_t)ds * a.n_params;
        const double tau = (KIND == KIND_TRIALWISE) ? prm[2] : prm[3];
        TrialF32 t;
        // a.dconst == nullptr: no prep_kernel ran in front of this launch, the constants are formed here
        const DsConst dc = a.dconst ? ((KIND == KIND_TRIALWISE) ? a.dconst[a.group[g]] : a.dconst[ds]) : prep_constants(prm, a.model, a.dt);
        if (KIND == KIND_TRIALWISE) trial_setup_trialwise(dc, (float)a.bound_in[g], t);
        else trial_setup_f32<(KIND == KIND_TRIALWISE ? KIND_FIXED : KIND)>(dc, trial_g, ds_g, a.key, t, cap);
        float x = t.x;
        uint32_t n = 0;
        bool alive = (fabsf(x) < t.h) && (a.max_steps > 0u);
        // software-pipelined: the normals of the next DDM_LATENCY_BLOCKS blocks are drawn while the current ones step (they
        // depend on the block index alone), so an iteration lasts max(generator chain, step chain) instead of their sum
        constexpr uint32_t NB = DDM_LATENCY_BLOCKS;
        Normals6Scaled z[NB];
        if (alive) {
#pragma unroll
            for (uint32_t j = 0; j < NB; j++) philox_pairs_lg2(j, trial_g, ds_g, STREAM_STEP, a.key, z[j]);
        }
        for (uint32_t blk = NB; alive; blk += NB) {
            Normals6Scaled y[NB];
#pragma unroll
            for (uint32_t j = 0; j < NB; j++) philox_pairs_lg2(blk + j, trial_g, ds_g, STREAM_STEP, a.key, y[j]);
#pragma unroll
            for (uint32_t j = 0; j < NB; j++) {
                spec_block(x, n, alive, t.c0, t.h, z[j]);
                alive = alive && (n < a.max_steps);
            }
#pragma unroll
            for (uint32_t j = 0; j < NB; j++) z[j] = y[j];
        }
        int choice = (x >= t.h) ? 1 : ((x <= -t.h) ? -1 : 0);
        if (n > a.max_steps) { n = a.max_steps; choice = 0; }  // whole blocks: a trial inside the boundaries after max_steps steps timed out
        const double ext = (KIND == KIND_TRIALWISE) ? a.bound_in[g] : (double)t.ext;
        double o0, o1;
        trial_outputs<(KIND == KIND_FIXED || KIND == KIND_DRIFT)>(a.flags, choice, n, a.dt, tau, ext, o0, o1);
        store_pair<OUT64>(a.out, g, o0, o1);
        if (a.steps_out) a.steps_out[g] = (int32_t)n;
        acc_steps = n;
        tout = (choice == 0);
        upper = (choice > 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_steps += __shfl_xor_sync(FULL_MASK, acc_steps, o);
        tout += __shfl_xor_sync(FULL_MASK, tout, o);
        upper += __shfl_xor_sync(FULL_MASK, upper, o);
        cap += __shfl_xor_sync(FULL_MASK, cap, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(a.stats + STAT_STEPS, acc_steps);
        if (tout) atomicAdd(a.stats + STAT_TIMEOUTS, (unsigned long long)tout);
        if (upper) atomicAdd(a.stats + STAT_UPPER, (unsigned long long)upper);
        if (cap) atomicAdd(a.stats + STAT_REJECT_CAP, (unsigned long long)cap);
    }
}

// --------------------------------------------------------------------------------
// general (two-latent, two-channel) model, one thread per trial: validation twin of
// persistent_kernel<KIND_GENERAL>
// --------------------------------------------------------------------------------
template <typename Real, bool BUFFER, bool OUT64>
__global__ void __launch_bounds__(128) general_generic_kernel(const RunArgs a, uint64_t total) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc_steps = 0;
    uint32_t tout = 0, upper = 0, cap = 0;
    if (g < total) {
        const uint32_t ds = (uint32_t)(g / a.n_trials);
        const uint32_t trial = (uint32_t)(g - (uint64_t)ds * a.n_trials);
        const uint32_t ds_g = ds + a.dataset_offset, trial_g = trial + a.trial_offset;
        const double *p = a.params + (size_t)ds * 24;
        uint32_t n = 0;
        int choice = 0;
        double e1 = 0.0, e2 = 0.0, final_ev = 0.0;
        if (sizeof(Real) == 4 && !BUFFER && !(a.flags & FLAG_REFERENCE_ARITHMETIC)) {
            // the persistent kernel's arithmetic, naive scheduling
            TrialF32 t;
            trial_setup_general(a.gconst[ds], trial_g, ds_g, a.key, t, cap);
            float x = t.x;
            uint32_t alive = ((fabsf(x) < t.h) && (a.max_steps > 0u)) ? 1u : 0u;
            for (uint32_t blk = 0; alive != 0u; blk++)
                step_block_f32<true>(blk, trial_g, ds_g, a.key, t, x, n, alive, a.max_steps);
            choice = (x >= t.h) ? 1 : ((x <= -t.h) ? -1 : 0);
            e1 = (double)t.ext;
            e2 = (double)t.ext2;
            final_ev = (double)__fmul_rn(__fadd_rn(x, t.h), t.u);
        } else {
            // the reference's formulas in Real, operation for operation (fp64: bit-equal to the numba loop)
            NormalStreamBuf buf{BUFFER ? a.dbg_z + a.dbg_off[g] : nullptr, BUFFER ? a.dbg_z + a.dbg_n : nullptr, false};
            double zc[6];
            uint32_t ztag = 0xffffffffu;
            auto normal = [&](uint32_t stream, uint32_t idx) -> double {
                if (BUFFER) return buf.next();
                const uint32_t b = idx / 6u, tag = b | (stream << 31);
                if (tag != ztag) {
                    if (sizeof(Real) == 8 && !(a.flags & 32)) {
                        philox_normals6_f64(b, trial_g, ds_g, stream, a.key, zc);
                    } else {
                        float zf[6];
                        philox_normals6_f32(b, trial_g, ds_g, stream, a.key, zf);
#pragma unroll
                        for (int i = 0; i < 6; i++) zc[i] = zf[i];
                    }
                    ztag = tag;
                }
                return zc[idx - 6u * b];
            };
            int ord = (int)p[20];
            if (ord < 0 || ord > 5) ord = 0;
            Real lat[3] = {(Real)p[0], (Real)p[2], (Real)p[4]};
            uint32_t cand[3] = {0u, 0u, 0u};
            for (int k = 0; k < 3; k++) {
                // permutations of (drift, boundary, dc): 012 021 102 120 201 210
                const int which = (k == 0) ? (ord >> 1) : ((k == 1) ? ((0x102021 >> (4 * ord)) & 3) : ((0x010212 >> (4 * ord)) & 3));
                const Real mu = (Real)p[2 * which], sd = (Real)p[2 * which + 1];
                if (sd == (Real)0) continue;
                for (;;) {
                    const uint32_t idx = (which == 0) ? 2u : (which == 1 ? 4u + 2u * cand[1] : 5u + 2u * cand[2]);
                    lat[which] = mu + sd * (Real)normal(STREAM_AUX, idx);
                    cand[which]++;
                    if (which == 0 || lat[which] > (Real)0) break;
                    if (cand[which] >= 3u * REJECT_CAP_BLOCKS - 2u) { cap++; lat[which] = (Real)1e-30; break; }
                }
            }
            const Real drift_t = lat[0], bound_t = lat[1], dc_t = lat[2];
            const Real dt = (Real)a.dt, sqrt_dt = (Real)a.sqrt_dt, beta = (Real)p[6];
            Real ev = bound_t * beta;
            while ((ev > (Real)0) && (ev < bound_t) && (n < a.max_steps)) {
                const Real z = (Real)normal(STREAM_STEP, n);
                const Real t1 = drift_t * dt;
                const Real t2 = sqrt_dt * dc_t;
                const Real t3 = t2 * z;
                ev = ev + (t1 + t3);
                n++;
            }
            choice = (ev >= bound_t) ? 1 : ((ev <= (Real)0) ? -1 : 0);
            final_ev = (double)ev;
            const int n_ext = (int)p[21];
            Real ext[2] = {(Real)0, (Real)0};
            for (int c = 0; c < 2 && c < n_ext; c++) {
                const double *e = p + 8 + 6 * c;
                const Real loc = ((Real)e[0] * drift_t + (Real)e[1] * bound_t) + (Real)e[2] * dc_t;
                const Real temp = loc + (Real)e[3] * (Real)normal(STREAM_AUX, (uint32_t)c);
                ext[c] = (temp - (Real)e[4]) / (Real)e[5];
            }
            e1 = (double)ext[0];
            e2 = (double)ext[1];
            if (BUFFER && buf.overrun) atomicAdd(a.stats + STAT_DBG_OVERRUN, 1ull);
        }
        double o0, o1;
        if ((int)p[22] == 0) {
            trial_outputs<true>(a.flags, choice, n, a.dt, p[7], 0.0, o0, o1);
            store_triple<OUT64>(a.out, g, o0, o1, (a.flags & 16) ? final_ev : e1);
        } else {
            trial_outputs<false>(a.flags, choice, n, a.dt, p[7], e1, o0, o1);
            store_triple<OUT64>(a.out, g, o0, o1, (a.flags & 16) ? final_ev : e2);
        }
        if (a.steps_out) a.steps_out[g] = (int32_t)n;
        acc_steps = n;
        tout = (choice == 0);
        upper = (choice > 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_steps += __shfl_xor_sync(FULL_MASK, acc_steps, o);
        tout += __shfl_xor_sync(FULL_MASK, tout, o);
        upper += __shfl_xor_sync(FULL_MASK, upper, o);
        cap += __shfl_xor_sync(FULL_MASK, cap, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(a.stats + STAT_STEPS, acc_steps);
        if (tout) atomicAdd(a.stats + STAT_TIMEOUTS, (unsigned long long)tout);
        if (upper) atomicAdd(a.stats + STAT_UPPER, (unsigned long long)upper);
        if (cap) atomicAdd(a.stats + STAT_REJECT_CAP, (unsigned long long)cap);
    }
}

// --------------------------------------------------------------------------------
// parity hooks
// --------------------------------------------------------------------------------
__global__ void export_normals_kernel(PhiloxKey key, uint32_t dataset, uint32_t trial, uint32_t stream,
                                      uint32_t first, uint32_t count, int f64, double *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t idx = first + i;
    const uint32_t b = idx / 6u, j = idx - 6u * b;
    if (f64) {
        double z[6];
        philox_normals6_f64(b, trial, dataset, stream, key, z);
        out[i] = z[j];
    } else {
        float z[6];
        philox_normals6_f32(b, trial, dataset, stream, key, z);
        out[i] = (double)z[j];
    }
}

__global__ void philox_blocks_kernel(const uint32_t *ctr, const uint32_t *key, uint32_t *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    philox4x32<10>(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], o);
    out[4 * i] = o[0]; out[4 * i + 1] = o[1]; out[4 * i + 2] = o[2]; out[4 * i + 3] = o[3];
}

// --------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------
cudaError_t launch_prep(const double *params, DsConst *dconst, uint32_t n_datasets, uint32_t n_params,
                        int model, double dt, cudaStream_t s) {
    const int block = 128;
    const unsigned grid = (n_datasets + block - 1) / block;
    prep_kernel<<<grid, block, 0, s>>>(params, dconst, n_datasets, n_params, model, dt);
    return cudaGetLastError();
}

template <int KIND>
static cudaError_t launch_persistent_kind(const RunArgs &a, bool out64, int grid, int block, cudaStream_t s) {
    if (out64) persistent_kernel<KIND, true><<<grid, block, 0, s>>>(a);
    else persistent_kernel<KIND, false><<<grid, block, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_persistent(const RunArgs &a, int kind, bool out64, int grid, int block, cudaStream_t s) {
    switch (kind) {
    case KIND_FIXED: return launch_persistent_kind<KIND_FIXED>(a, out64, grid, block, s);
    case KIND_BOUND: return launch_persistent_kind<KIND_BOUND>(a, out64, grid, block, s);
    case KIND_DC: return launch_persistent_kind<KIND_DC>(a, out64, grid, block, s);
    case KIND_DRIFT: return launch_persistent_kind<KIND_DRIFT>(a, out64, grid, block, s);
    case KIND_GENERAL: return launch_persistent_kind<KIND_GENERAL>(a, out64, grid, block, s);
    default: return cudaErrorInvalidValue;
    }
}

template <int KIND>
static cudaError_t launch_tile_kind(const RunArgs &a, bool out64, int grid, int block, size_t smem, cudaStream_t s) {
    // the tile buffers want most of the SM's 228 KB as shared memory (6 blocks x up to 28 KB)
    static bool configured[2] = {false, false};
    if (!configured[out64 ? 1 : 0]) {
        cudaError_t e = out64 ? cudaFuncSetAttribute(tile_kernel<KIND, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                     cudaSharedmemCarveoutMaxShared)
                              : cudaFuncSetAttribute(tile_kernel<KIND, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                     cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured[out64 ? 1 : 0] = true;
    }
    if (out64) tile_kernel<KIND, true><<<grid, block, smem, s>>>(a);
    else tile_kernel<KIND, false><<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_tile(const RunArgs &a, int kind, bool out64, int grid, int block, size_t smem, cudaStream_t s) {
    switch (kind) {
    case KIND_FIXED: return launch_tile_kind<KIND_FIXED>(a, out64, grid, block, smem, s);
    case KIND_BOUND: return launch_tile_kind<KIND_BOUND>(a, out64, grid, block, smem, s);
    case KIND_DC: return launch_tile_kind<KIND_DC>(a, out64, grid, block, smem, s);
    case KIND_TRIALWISE: return launch_tile_kind<KIND_TRIALWISE>(a, out64, grid, block, smem, s);
    case KIND_DRIFT: return launch_tile_kind<KIND_DRIFT>(a, out64, grid, block, smem, s);
    case KIND_GENERAL: return launch_tile_kind<KIND_GENERAL>(a, out64, grid, block, smem, s);
    default: return cudaErrorInvalidValue;
    }
}

template <int KIND>
static size_t tile_smem_of(int block) {
    return (size_t)(block / 32) * ((size_t)TileLayout<KIND>::WORDS * TileLayout<KIND>::T + M_WORDS) * sizeof(uint32_t);
}

size_t tile_kernel_smem_bytes(int kind, int block) {
    switch (kind) {
    case KIND_FIXED: return tile_smem_of<KIND_FIXED>(block);
    case KIND_BOUND: return tile_smem_of<KIND_BOUND>(block);
    case KIND_DC: return tile_smem_of<KIND_DC>(block);
    case KIND_TRIALWISE: return tile_smem_of<KIND_TRIALWISE>(block);
    case KIND_DRIFT: return tile_smem_of<KIND_DRIFT>(block);
    default: return tile_smem_of<KIND_GENERAL>(block);
    }
}

uint32_t tile_kernel_max_tile(int kind) { return kind == KIND_GENERAL ? TileLayout<KIND_GENERAL>::T : 128u; }

int tile_kernel_max_blocks_per_sm(int kind, bool out64, int block, size_t smem) {
    int nb = 0;
    cudaError_t e = cudaErrorInvalidValue;
#define DDM_OCC(K, O) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tile_kernel<K, O>, block, smem)
    if (kind == KIND_FIXED) { if (out64) DDM_OCC(KIND_FIXED, true); else DDM_OCC(KIND_FIXED, false); }
    else if (kind == KIND_BOUND) { if (out64) DDM_OCC(KIND_BOUND, true); else DDM_OCC(KIND_BOUND, false); }
    else if (kind == KIND_DC) { if (out64) DDM_OCC(KIND_DC, true); else DDM_OCC(KIND_DC, false); }
    else if (kind == KIND_TRIALWISE) { if (out64) DDM_OCC(KIND_TRIALWISE, true); else DDM_OCC(KIND_TRIALWISE, false); }
    else if (kind == KIND_DRIFT) { if (out64) DDM_OCC(KIND_DRIFT, true); else DDM_OCC(KIND_DRIFT, false); }
    else if (kind == KIND_GENERAL) { if (out64) DDM_OCC(KIND_GENERAL, true); else DDM_OCC(KIND_GENERAL, false); }
#undef DDM_OCC
    return (e == cudaSuccess) ? nb : -1;
}

int persistent_block_size() { return DDM_PERSISTENT_BLOCK; }

int persistent_max_blocks_per_sm(int kind, bool out64, int block) {
    int nb = 0;
    cudaError_t e = cudaErrorInvalidValue;
#define DDM_OCC(K, O) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, persistent_kernel<K, O>, block, 0)
    if (kind == KIND_FIXED) { if (out64) DDM_OCC(KIND_FIXED, true); else DDM_OCC(KIND_FIXED, false); }
    else if (kind == KIND_BOUND) { if (out64) DDM_OCC(KIND_BOUND, true); else DDM_OCC(KIND_BOUND, false); }
    else if (kind == KIND_DC) { if (out64) DDM_OCC(KIND_DC, true); else DDM_OCC(KIND_DC, false); }
    else if (kind == KIND_DRIFT) { if (out64) DDM_OCC(KIND_DRIFT, true); else DDM_OCC(KIND_DRIFT, false); }
    else if (kind == KIND_GENERAL) { if (out64) DDM_OCC(KIND_GENERAL, true); else DDM_OCC(KIND_GENERAL, false); }
#undef DDM_OCC
    return (e == cudaSuccess) ? nb : -1;
}

template <typename Real, int KIND, bool BUFFER>
static cudaError_t launch_generic_3(const RunArgs &a, bool out64, uint64_t total, cudaStream_t s) {
    const int block = 128;
    const unsigned grid = (unsigned)((total + block - 1) / block);
    if (grid == 0) return cudaSuccess;
    if (out64) generic_kernel<Real, KIND, BUFFER, true><<<grid, block, 0, s>>>(a, total);
    else generic_kernel<Real, KIND, BUFFER, false><<<grid, block, 0, s>>>(a, total);
    return cudaGetLastError();
}

template <typename Real, int KIND>
static cudaError_t launch_generic_2(const RunArgs &a, bool buffer_src, bool out64, uint64_t total, cudaStream_t s) {
    return buffer_src ? launch_generic_3<Real, KIND, true>(a, out64, total, s)
                      : launch_generic_3<Real, KIND, false>(a, out64, total, s);
}

template <typename Real>
static cudaError_t launch_generic_1(const RunArgs &a, int kind, bool buffer_src, bool out64, uint64_t total,
                                    cudaStream_t s) {
    switch (kind) {
    case KIND_FIXED: return launch_generic_2<Real, KIND_FIXED>(a, buffer_src, out64, total, s);
    case KIND_BOUND: return launch_generic_2<Real, KIND_BOUND>(a, buffer_src, out64, total, s);
    case KIND_DC: return launch_generic_2<Real, KIND_DC>(a, buffer_src, out64, total, s);
    case KIND_TRIALWISE: return launch_generic_2<Real, KIND_TRIALWISE>(a, buffer_src, out64, total, s);
    case KIND_DRIFT: return launch_generic_2<Real, KIND_DRIFT>(a, buffer_src, out64, total, s);
    default: return cudaErrorInvalidValue;
    }
}

template <int KIND>
static cudaError_t launch_latency_kind(const RunArgs &a, bool out64, uint64_t total, cudaStream_t s) {
    const unsigned grid = (unsigned)((total + DDM_LATENCY_THREADS - 1) / DDM_LATENCY_THREADS);
    if (grid == 0) return cudaSuccess;
    if (out64) latency_kernel<KIND, true><<<grid, DDM_LATENCY_THREADS, 0, s>>>(a, total);
    else latency_kernel<KIND, false><<<grid, DDM_LATENCY_THREADS, 0, s>>>(a, total);
    return cudaGetLastError();
}

cudaError_t launch_latency(const RunArgs &a, int kind, bool out64, uint64_t total_trials, cudaStream_t s) {
    switch (kind) {
    case KIND_FIXED: return launch_latency_kind<KIND_FIXED>(a, out64, total_trials, s);
    case KIND_BOUND: return launch_latency_kind<KIND_BOUND>(a, out64, total_trials, s);
    case KIND_DC: return launch_latency_kind<KIND_DC>(a, out64, total_trials, s);
    case KIND_TRIALWISE: return launch_latency_kind<KIND_TRIALWISE>(a, out64, total_trials, s);
    case KIND_DRIFT: return launch_latency_kind<KIND_DRIFT>(a, out64, total_trials, s);
    default: return cudaErrorInvalidValue;
    }
}

template <typename Real>
static cudaError_t launch_general_generic(const RunArgs &a, bool buffer_src, bool out64, uint64_t total, cudaStream_t s) {
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (grid == 0) return cudaSuccess;
    if (buffer_src) {
        if (out64) general_generic_kernel<Real, true, true><<<grid, 128, 0, s>>>(a, total);
        else general_generic_kernel<Real, true, false><<<grid, 128, 0, s>>>(a, total);
    } else {
        if (out64) general_generic_kernel<Real, false, true><<<grid, 128, 0, s>>>(a, total);
        else general_generic_kernel<Real, false, false><<<grid, 128, 0, s>>>(a, total);
    }
    return cudaGetLastError();
}

cudaError_t launch_prep_general(const double *params, GenConst *gconst, uint32_t n_datasets, double dt, cudaStream_t s) {
    if (n_datasets == 0) return cudaSuccess;
    prep_general_kernel<<<(n_datasets + 127) / 128, 128, 0, s>>>(params, gconst, n_datasets, dt);
    return cudaGetLastError();
}

cudaError_t launch_generic(const RunArgs &a, int kind, bool f64, bool buffer_src, bool out64,
                           uint64_t total_trials, cudaStream_t s) {
    if (kind == KIND_GENERAL)
        return f64 ? launch_general_generic<double>(a, buffer_src, out64, total_trials, s)
                   : launch_general_generic<float>(a, buffer_src, out64, total_trials, s);
    return f64 ? launch_generic_1<double>(a, kind, buffer_src, out64, total_trials, s)
               : launch_generic_1<float>(a, kind, buffer_src, out64, total_trials, s);
}

cudaError_t launch_export_normals(PhiloxKey key, uint32_t dataset, uint32_t trial, uint32_t stream,
                                  uint32_t first, uint32_t count, bool f64, double *out_dev, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    const int block = 128;
    export_normals_kernel<<<(count + block - 1) / block, block, 0, s>>>(key, dataset, trial, stream, first,
                                                                         count, f64 ? 1 : 0, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_philox_blocks(const uint32_t *ctr, const uint32_t *key, uint32_t *out, int64_t n,
                                 cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int block = 128;
    philox_blocks_kernel<<<(unsigned)((n + block - 1) / block), block, 0, s>>>(ctr, key, out, n);
    return cudaGetLastError();
}

}  // namespace ddm
