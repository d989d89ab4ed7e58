// ddm_evidence.cu -- evidence-path variants: (rt, choice, path[n_obs]) per trial.
//
// Replaces (reference, retired model zoo):
//   retired_models/basic_ddm_dc_evidence.py:87-151        200 obs, noise sigma1, per-trial z-score   (mode 1)
//   retired_models/basic_ddm_dc_evidence2.py:83-150       200 obs, noise sigma1, dataset-level
//                                                          (x - mean(path_means)) / std(path_means) (mode 2)
//   retired_models/basic_ddm_dc_evidence_no_noise2.py:82-147  400 obs, noise .001, per-trial z-score (mode 1)
// params[6] = drift, boundary, beta, tau, dc, sigma1.  path[k] = evidence after Euler step k+1 for
// k < n, held at the final evidence for k >= n, plus sigma1 * z_noise[k]; then standardised.
//
//   production (fp32)         two kernels.  (1) record_kernel steps the trials (one lane per trial, persistent warps,
//                             the Philox counters and step arithmetic of DDM_MODEL_BASIC) and stores the first
//                             n_obs states of each trial one whole 32-byte sector at a time straight from
//                             registers, plus 8 bytes per trial (steps, choice, final state).
//                             (2) evidence_post_kernel: eight lanes per trial turn the recorded states into the
//                             observed path -- evidence units, held at the final evidence after the crossing, noise
//                             normals from the aux Philox stream (six per lane), mean / variance by shuffle
//                             reduction -- and write the row with coalesced stores.  800 B/trial of output make
//                             this the one DDM path where stores matter.
//   evidence_generic_kernel   validation (fp64): one thread per trial, the reference's operation order and
//                             left-to-right sums (numba's array_mean / array_var); shared-increment mode.
//   dataset_stats / finalize  mode 2's second pass, and the dtype conversion of the validation path.
#include "ddm_kernels.cuh"

namespace ddm {


// --------------------------------------------------------------------------------------------
// production, first kernel: step and record
// --------------------------------------------------------------------------------------------
// One lane per trial, persistent warps that claim (dataset, tile) work items from a global counter, like the
// simulator's own kernels -- but time runs in *periods* of four Philox blocks = 24 Euler steps = three 32-byte
// sectors of a trial's recorded path, and lanes are handed new trials only between periods.  Every lane of a warp
// is then at the same phase of its trial's 24-step group, so the group's states can stay in registers under
// static names and leave as whole sectors: block 1 completes sector 0 (states 0..7), block 2 sector 1, block 3
// sector 2 -- two STG.128 per sector and lane, no staging in shared memory, no per-lane ring position, no
// partial-sector writes (8-byte stores straight from every block made the kernel L2-bound; a per-lane
// shared-memory ring, round 1's fix, doubled the instruction count and capped occupancy at 38 %).
// A lane whose trial ends inside a period goes on "recording" its frozen state to the end of the period: those
// positions lie past the trial's last step and the post kernel never reads them (at most two surplus sectors per
// trial, instead of a tail-flush path in every block).  A finished lane waits two blocks on average for the
// period to end -- 5 % of a 250-step trial, the price of the aligned phases.
// Philox counters, set-up and step arithmetic are those of DDM_MODEL_BASIC's kernels: a trial's steps, choice and
// states do not depend on which kernel ran it.
//   VEC: n_obs is a multiple of 8 (rows are sector-aligned); otherwise guarded scalar stores, six per block.
constexpr int RECORD_BLOCK = 256;
#ifndef DDM_RECORD_MIN_BLOCKS
#define DDM_RECORD_MIN_BLOCKS 5
#endif
template <bool VEC>
__global__ void __launch_bounds__(RECORD_BLOCK, DDM_RECORD_MIN_BLOCKS) record_kernel(const __grid_constant__ RunArgs a) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const RngConsts rk = pinned_rng_consts();

    // warp-uniform tile cursor and the tile's dataset constants (drift*dt/U, start state, half-width)
    uint32_t cur = 0, end = 0, tile_ds = 0;
    bool more = true;
    float t_c0 = 0.f, t_x0 = 0.f, t_h = 0.f;

    // per-lane trial
    float x = 0.f, h = 0.f, c0 = 0.f;
    uint32_t n = 0, per = 0, trial = 0, ds = 0;
    uint32_t p = 0;    // 1 = stepping
    bool has = false;  // holds a trial (stepping, or finished and waiting to be emitted)

    unsigned long long acc_steps = 0;
    uint32_t acc_timeouts = 0, acc_upper = 0;
    uint2 *meta = reinterpret_cast<uint2 *>(a.out);
    const int thr = a.refill_threshold > 1 ? a.refill_threshold : 1;

    for (;;) {
        const unsigned live = __ballot_sync(FULL_MASK, has && p != 0u);
        const bool work_left = more || cur != end;
        if (__popc(~live) >= (work_left ? thr : 32)) {
            // ---- finish: 8 bytes per trial ------------------------------------------------------
            if (has && p == 0u) {
                int choice = (x >= h) ? 1 : ((x <= -h) ? -1 : 0);
                // lanes run whole 6-step blocks: a trial still inside the boundaries after max_steps steps is a
                // timeout whatever it did in the surplus steps of its last block
                if (n > a.max_steps) { n = a.max_steps; choice = 0; }
                const uint64_t idx = (uint64_t)ds * a.n_trials + trial;
                if (DDM_CHECK(a.stats, ds < a.n_datasets && trial < a.n_trials))
                    meta[idx] = make_uint2((uint32_t)wire_pack(n, choice), __float_as_uint(x));
                acc_steps += n;
                acc_timeouts += (choice == 0);
                acc_upper += (choice > 0);
                has = false;
            }
            // ---- refill: hand out trials of the current tile, claiming tiles as needed ----------
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
                    if (a.tiles_per_dataset == 1u) {
                        tile_ds = (uint32_t)w;
                    } else {
                        tile_ds = (uint32_t)w / a.tiles_per_dataset;  // host keeps n_items < 2^32
                        ti = (uint32_t)w - tile_ds * a.tiles_per_dataset;
                    }
                    cur = ti * a.tile;
                    end = min(cur + a.tile, a.n_trials);
                    const float4 c = __ldg(reinterpret_cast<const float4 *>(a.dconst + tile_ds));
                    t_c0 = c.x; t_x0 = c.y; t_h = c.z;
                }
                const uint32_t rank = __popc(empty & lt_mask);
                const uint32_t avail = end - cur;
                if (!has && rank < avail) {
                    ds = tile_ds;
                    trial = cur + rank;
                    x = t_x0; h = t_h; c0 = t_c0;
                    n = 0;
                    per = 0;
                    has = true;
                    p = ((fabsf(x) < h) && (a.max_steps > 0u)) ? 1u : 0u;
                }
                cur += min((uint32_t)__popc(empty), avail);
            }
            if (!__any_sync(FULL_MASK, has)) break;
        }

        // ---- one period: four blocks, three sectors ---------------------------------------------
        const uint32_t n0 = 24u * per;  // steps a lane that is still stepping has taken
        const bool rec = has && p != 0u && n0 < a.n_obs;
        float *dst = a.rec_path + ((uint64_t)ds * a.n_trials + trial) * a.n_obs + n0;
        const uint32_t tg = trial + a.trial_offset, dg = ds + a.dataset_offset, b0 = 4u * per;
        Normals6Scaled z;
        float r[6];
        if (VEC) {
            float k0, k1, k2, k3, k4, k5;  // states held over from the previous block
            philox_pairs_lg2(b0, tg, dg, STREAM_STEP, a.key, rk, z);
            euler6_rec(x, n, p, c0, h, z, a.max_steps, r);
            k0 = r[0]; k1 = r[1]; k2 = r[2]; k3 = r[3]; k4 = r[4]; k5 = r[5];
            philox_pairs_lg2(b0 + 1u, tg, dg, STREAM_STEP, a.key, rk, z);
            euler6_rec(x, n, p, c0, h, z, a.max_steps, r);
            if (rec) {  // states 0..7 (n0 < n_obs and both are multiples of 8: the sector lies inside the row)
                reinterpret_cast<float4 *>(dst)[0] = make_float4(k0, k1, k2, k3);
                reinterpret_cast<float4 *>(dst)[1] = make_float4(k4, k5, r[0], r[1]);
            }
            k0 = r[2]; k1 = r[3]; k2 = r[4]; k3 = r[5];
            philox_pairs_lg2(b0 + 2u, tg, dg, STREAM_STEP, a.key, rk, z);
            euler6_rec(x, n, p, c0, h, z, a.max_steps, r);
            if (rec && n0 + 8u < a.n_obs) {  // states 8..15
                reinterpret_cast<float4 *>(dst)[2] = make_float4(k0, k1, k2, k3);
                reinterpret_cast<float4 *>(dst)[3] = make_float4(r[0], r[1], r[2], r[3]);
            }
            k0 = r[4]; k1 = r[5];
            philox_pairs_lg2(b0 + 3u, tg, dg, STREAM_STEP, a.key, rk, z);
            euler6_rec(x, n, p, c0, h, z, a.max_steps, r);
            if (rec && n0 + 16u < a.n_obs) {  // states 16..23
                reinterpret_cast<float4 *>(dst)[4] = make_float4(k0, k1, r[0], r[1]);
                reinterpret_cast<float4 *>(dst)[5] = make_float4(r[2], r[3], r[4], r[5]);
            }
        } else {
#pragma unroll 1
            for (uint32_t ph = 0; ph < 4u; ph++) {
                philox_pairs_lg2(b0 + ph, tg, dg, STREAM_STEP, a.key, rk, z);
                euler6_rec(x, n, p, c0, h, z, a.max_steps, r);
#pragma unroll
                for (uint32_t i = 0; i < 6u; i++)
                    if (rec && n0 + 6u * ph + i < a.n_obs) dst[6u * ph + i] = r[i];
            }
        }
        per++;
    }

    // ---- per-warp statistics ----------------------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_steps += __shfl_xor_sync(FULL_MASK, acc_steps, o);
        acc_timeouts += __shfl_xor_sync(FULL_MASK, acc_timeouts, o);
        acc_upper += __shfl_xor_sync(FULL_MASK, acc_upper, o);
    }
    if (lane == 0) {
        atomicAdd(a.stats + STAT_STEPS, acc_steps);
        atomicAdd(a.stats + STAT_TIMEOUTS, (unsigned long long)acc_timeouts);
        atomicAdd(a.stats + STAT_UPPER, (unsigned long long)acc_upper);
    }
}

cudaError_t launch_record(const RunArgs &a, int grid, int block, cudaStream_t s) {
    if ((a.n_obs & 7u) == 0u) record_kernel<true><<<grid, block, 0, s>>>(a);
    else record_kernel<false><<<grid, block, 0, s>>>(a);
    return cudaGetLastError();
}

int record_block_size() { return RECORD_BLOCK; }

int record_max_blocks_per_sm(int block) {
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, record_kernel<true>, block, 0);
    return (e == cudaSuccess) ? nb : -1;
}

// --------------------------------------------------------------------------------------------
// production, second kernel: a warp per trial finishes the recorded path
// --------------------------------------------------------------------------------------------
// G lanes work on one trial, 32 / G trials per warp at a time.  A trial's 200 noise normals are 34 Philox blocks:
// on 32 lanes that is two passes with the second nearly empty; on 8 lanes it is five passes for four trials
// (1.25 per trial instead of 2), and the reductions are three shuffle steps instead of five.
template <bool OUT64, int G>
__global__ void __launch_bounds__(256) evidence_post_kernel(const EvidenceArgs a, uint64_t total) {
    // Per-trial staging row: every global access below is contiguous over the G lanes (k = sl, sl + G, ...: whole
    // 32-byte sectors); the noise normals, which come six per Philox block and per lane, meet the row in shared
    // memory.
    extern __shared__ float post_smem[];
    constexpr unsigned TPW = 32u / G;  // trials per warp
    const unsigned lane = threadIdx.x & 31u, sl = lane & (G - 1u), sub = lane / G;
    float *s = post_smem + ((size_t)(threadIdx.x >> 5) * TPW + sub) * a.n_obs;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t cols = 2u + a.n_obs;
    const uint32_t n_blocks = (a.n_obs + 5u) / 6u;
    const float inv_n = 1.f / (float)a.n_obs;
    const bool small = total <= 0xffffffffull;  // 32-bit index arithmetic when the batch allows
    for (uint64_t base = warp0 * TPW; base < total; base += n_warps * TPW) {
        const bool valid = base + sub < total;  // the last warp may hold fewer than TPW trials: idle groups redo the
        const uint64_t g = valid ? base + sub : total - 1;  // last trial and store nothing
        uint32_t ds, trial;
        if (small) {
            ds = (uint32_t)g / a.n_trials;
            trial = (uint32_t)g - ds * a.n_trials;
        } else {
            ds = (uint32_t)(g / a.n_trials);
            trial = (uint32_t)(g - (uint64_t)ds * a.n_trials);
        }
        const uint2 mt = a.rec_meta[g];  // ((steps << 2) | (choice + 1), final state)
        const uint32_t nj = mt.x >> 2;
        const float h = a.dconst[ds].v[2], u = a.dconst[ds].v[3];
        const float sigma1 = (float)a.params[(size_t)ds * 6 + 5];
        const float evj = __fmul_rn(__fadd_rn(__uint_as_float(mt.y), h), u);
        const float *row_in = a.rec_path + g * a.n_obs;
        // 1. recorded states -> evidence units, held at the final evidence after the crossing
        for (uint32_t k = sl; k < a.n_obs; k += G)
            s[k] = (k < nj) ? __fmul_rn(__fadd_rn(row_in[k], h), u) : evj;
        __syncwarp();
        // 2. + sigma1 * z_noise[k]: lane owns Philox blocks sl, sl + G, ... of the trial's aux stream;
        //    the row sum is taken on the way
        float sum = 0.f;
        for (uint32_t b = sl; b < n_blocks; b += G) {
            float z[6];
            philox_normals6_f32(b, trial + a.trial_offset, ds + a.dataset_offset, STREAM_AUX, a.key, z);
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const uint32_t k = 6u * b + i;
                if (k < a.n_obs) {
                    const float v = __fmaf_rn(sigma1, z[i], s[k]);
                    s[k] = v;
                    sum += v;
                }
            }
        }
        __syncwarp();
        // 3. mean (and variance) over the row
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, o);
        const float mean = sum * inv_n;
        float scale = 1.f, shift = 0.f;
        if (a.mode == 1) {
            float ssd = 0.f;
            for (uint32_t k = sl; k < a.n_obs; k += G) {
                const float d = s[k] - mean;
                ssd = __fmaf_rn(d, d, ssd);
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) ssd += __shfl_xor_sync(FULL_MASK, ssd, o);
            scale = 1.f / sqrtf(ssd * inv_n);
            shift = mean;
        }
        // 4. the row
        const uint64_t row = g * cols;
        if (sl == 0 && valid) {
            double rt, ch;  // (rt, choice) in the reference's fp64 arithmetic (basic_ddm_dc_evidence.py:127-135)
            trial_outputs<true>(0, (int)(mt.x & 3u) - 1, nj, a.dt, a.params[(size_t)ds * 6 + 3], 0.0, rt, ch);
            if (OUT64) {
                double *o = reinterpret_cast<double *>(a.out) + row;
                o[0] = rt;
                o[1] = ch;
            } else {
                float *o = reinterpret_cast<float *>(a.out) + row;
                o[0] = (float)rt;
                o[1] = (float)ch;
            }
            if (a.mode == 2) a.path_means[g] = (double)mean;
        }
        if (valid) {
            for (uint32_t k = sl; k < a.n_obs; k += G) {
                const float v = (s[k] - shift) * scale;
                if (OUT64) reinterpret_cast<double *>(a.out)[row + 2 + k] = (double)v;
                else reinterpret_cast<float *>(a.out)[row + 2 + k] = v;
            }
        }
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------------
// validation: one thread per trial, fp64, reference operation order; rows go to an fp64 scratch
// --------------------------------------------------------------------------------------------
template <bool BUFFER>
__global__ void __launch_bounds__(128) evidence_generic_kernel(const EvidenceArgs a, uint64_t total) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long steps = 0;
    uint32_t tout = 0, upper = 0;
    if (g < total) {
        const uint32_t ds = (uint32_t)(g / a.n_trials);
        const uint32_t trial = (uint32_t)(g - (uint64_t)ds * a.n_trials);
        const uint32_t ds_g = ds + a.dataset_offset, trial_g = trial + a.trial_offset;
        const double *prm = a.params + (size_t)ds * 6;
        const double drift = prm[0], boundary = prm[1], beta = prm[2], tau = prm[3], dcoef = prm[4], sigma1 = prm[5];
        double *row = a.scratch + g * (2ull + a.n_obs);
        double *path = row + 2;
        const double *zbuf = BUFFER ? a.dbg_z + a.dbg_off[g] : nullptr;
        const double *zend = BUFFER ? a.dbg_z + a.dbg_n : nullptr;
        bool overrun = false;
        double zc[6];
        uint32_t ztag = 0xffffffffu;
        auto normal = [&](uint32_t stream, uint32_t idx) -> double {
            if (BUFFER) {
                if (zbuf >= zend) { overrun = true; return 0.0; }
                return *zbuf++;
            }
            const uint32_t b = idx / 6u, tag = b | (stream << 31);
            if (tag != ztag) {
                philox_normals6_f64(b, trial_g, ds_g, stream, a.key, zc);
                ztag = tag;
            }
            return zc[idx - 6u * b];
        };
        for (uint32_t k = 0; k < a.n_obs; k++) path[k] = 0.0;
        uint32_t n = 0;
        double ev = boundary * beta;
        while ((ev > 0.0) && (ev < boundary) && (n < a.max_steps)) {
            const double z = normal(STREAM_STEP, n);
            const double t1 = drift * a.dt;
            const double t2 = a.sqrt_dt * dcoef;
            const double t3 = t2 * z;
            ev = ev + (t1 + t3);
            if (n < a.n_obs) path[n] = ev;
            n++;
        }
        for (uint32_t k = n; k < a.n_obs; k++) path[k] = ev;
        for (uint32_t k = 0; k < a.n_obs; k++) path[k] = path[k] + (0.0 + sigma1 * normal(STREAM_AUX, k));
        double c = 0.0;
        for (uint32_t k = 0; k < a.n_obs; k++) c += path[k];
        const double mean = c / (double)a.n_obs;
        if (a.mode == 1) {
            double ssd = 0.0;
            for (uint32_t k = 0; k < a.n_obs; k++) {
                const double d = path[k] - mean;
                ssd += d * d;
            }
            const double sd = sqrt(ssd / (double)a.n_obs);
            for (uint32_t k = 0; k < a.n_obs; k++) path[k] = (path[k] - mean) / sd;
        } else if (a.mode == 2) {
            a.path_means[g] = mean;
        }
        const int choice = (ev >= boundary) ? 1 : ((ev <= 0.0) ? -1 : 0);
        row[0] = __dadd_rn(__dmul_rn((double)n, a.dt), tau);
        row[1] = (double)choice;
        steps = n;
        tout = (choice == 0);
        upper = (choice > 0);
        if (BUFFER && overrun) atomicAdd(a.stats + STAT_DBG_OVERRUN, 1ull);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_xor_sync(FULL_MASK, steps, o);
        tout += __shfl_xor_sync(FULL_MASK, tout, o);
        upper += __shfl_xor_sync(FULL_MASK, upper, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(a.stats + STAT_STEPS, steps);
        if (tout) atomicAdd(a.stats + STAT_TIMEOUTS, (unsigned long long)tout);
        if (upper) atomicAdd(a.stats + STAT_UPPER, (unsigned long long)upper);
    }
}

// mode 2: mean and std of a dataset's per-trial path means, left to right (numba's array_mean / array_var)
__global__ void evidence_dataset_stats_kernel(const double *__restrict__ path_means, double *__restrict__ ds_stats,
                                              uint32_t n_datasets, uint32_t n_trials) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_datasets) return;
    const double *v = path_means + (size_t)d * n_trials;
    double c = 0.0;
    for (uint32_t i = 0; i < n_trials; i++) c += v[i];
    const double m = c / (double)n_trials;
    double ssd = 0.0;
    for (uint32_t i = 0; i < n_trials; i++) {
        const double dd = v[i] - m;
        ssd += dd * dd;
    }
    ds_stats[2 * d] = m;
    ds_stats[2 * d + 1] = sqrt(ssd / (double)n_trials);
}

// scratch (fp64 rows) or the output itself (in place) -> output dtype, applying mode 2's dataset-level
// standardisation to the path columns
template <typename Src, typename Dst>
// src and dst may be the same buffer (mode 2 standardises the production rows in place: each thread reads and
// writes its own element only), so neither is __restrict__.
__global__ void evidence_finalize_kernel(const Src *src, Dst *dst,
                                         const double *__restrict__ ds_stats, uint64_t total, uint32_t cols,
                                         uint32_t n_trials, int standardize) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint64_t row = i / cols;
    const uint32_t col = (uint32_t)(i - row * cols);
    double v = (double)src[i];
    if (standardize && col >= 2u) {
        const uint64_t d = row / n_trials;
        v = (v - ds_stats[2 * d]) / ds_stats[2 * d + 1];
    }
    dst[i] = (Dst)v;
}

// --------------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------------
template <int G>
static cudaError_t launch_evidence_post_g(const EvidenceArgs &a, bool out64, uint64_t total, int sm_count, cudaStream_t s) {
    constexpr uint64_t per_block = 8ull * (32 / G);  // 8 warps per block, 32 / G trials per warp per pass
    uint64_t grid = (total + per_block - 1) / per_block;
    const uint64_t cap = (uint64_t)sm_count * 8 * 4;
    if (grid > cap) grid = cap;
    const size_t smem = (size_t)per_block * a.n_obs * sizeof(float);
    if (out64) evidence_post_kernel<true, G><<<(unsigned)grid, 256, smem, s>>>(a, total);
    else evidence_post_kernel<false, G><<<(unsigned)grid, 256, smem, s>>>(a, total);
    return cudaGetLastError();
}

cudaError_t launch_evidence_post(const EvidenceArgs &a, bool out64, uint64_t total, int sm_count, cudaStream_t s) {
    if (total == 0) return cudaSuccess;
    // the widest split whose staging rows fit the 48 KB a block gets without opting in: 8 lanes per trial up to
    // 384 observations (the reference uses 200 and 400), then 16, then the whole warp
    const size_t row = (size_t)a.n_obs * sizeof(float);
    if (32 * row <= 48 * 1024) return launch_evidence_post_g<8>(a, out64, total, sm_count, s);
    if (16 * row <= 48 * 1024) return launch_evidence_post_g<16>(a, out64, total, sm_count, s);
    return launch_evidence_post_g<32>(a, out64, total, sm_count, s);
}

cudaError_t launch_evidence_generic(const EvidenceArgs &a, bool buffer_src, uint64_t total, cudaStream_t s) {
    if (total == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (buffer_src) evidence_generic_kernel<true><<<grid, 128, 0, s>>>(a, total);
    else evidence_generic_kernel<false><<<grid, 128, 0, s>>>(a, total);
    return cudaGetLastError();
}

cudaError_t launch_evidence_dataset_stats(const double *path_means, double *ds_stats, uint32_t n_datasets,
                                          uint32_t n_trials, cudaStream_t s) {
    if (n_datasets == 0) return cudaSuccess;
    evidence_dataset_stats_kernel<<<(n_datasets + 63) / 64, 64, 0, s>>>(path_means, ds_stats, n_datasets, n_trials);
    return cudaGetLastError();
}

cudaError_t launch_evidence_finalize(const void *src, bool src64, void *dst, bool dst64, const double *ds_stats,
                                     uint64_t total, uint32_t cols, uint32_t n_trials, bool standardize, cudaStream_t s) {
    if (total == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total + 255) / 256);
    const int st = standardize ? 1 : 0;
    if (src64 && dst64)
        evidence_finalize_kernel<double, double><<<grid, 256, 0, s>>>((const double *)src, (double *)dst, ds_stats, total, cols, n_trials, st);
    else if (src64 && !dst64)
        evidence_finalize_kernel<double, float><<<grid, 256, 0, s>>>((const double *)src, (float *)dst, ds_stats, total, cols, n_trials, st);
    else if (!src64 && !dst64)
        evidence_finalize_kernel<float, float><<<grid, 256, 0, s>>>((const float *)src, (float *)dst, ds_stats, total, cols, n_trials, st);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace ddm
