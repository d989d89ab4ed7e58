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
//   VEC: n_obs is a multiple of 8 (rows are sector-aligned); otherwise guarded scalar stores, six per block, rows in
//   natural order.
// Layout of a VEC row (what the post kernel wants to read): the post kernel gives a trial G lanes, lane sl owning the
// six observations of Philox block b = sl + G it -- 24 bytes at a 24-byte stride across lanes, a third of every sector
// per load instruction if the row were in time order.  The row is therefore stored in chunks of 6 G observations
// (= G blocks), each chunk as three planes of G pairs: observation 6 b + 2 j + e of chunk it sits at
// 6 G it + 2 G j + 2 (b mod G) + e, so that the post kernel's j-th load is 8 contiguous bytes per lane, G lanes
// contiguous.  For this kernel nothing changes but which register goes where: a period's four blocks are four
// neighbouring lanes' worth of one chunk, and plane j of the period -- pairs (2j, 2j+1) of the four blocks -- is
// again one aligned 32-byte sector.  Rows are padded to whole chunks (a.rec_stride floats).
constexpr int RECORD_BLOCK = 256;
// four resident blocks (64 registers: a period's 24 states stay in registers until its three sectors are written);
// at five (48 registers) ptxas interleaves the stores with the next blocks' arithmetic and the kernel waits on the
// store queue: 1.32 instead of 0.93 ms per 2M trials
#ifndef DDM_RECORD_MIN_BLOCKS
#define DDM_RECORD_MIN_BLOCKS 4
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
        const uint32_t tg = trial + a.trial_offset, dg = ds + a.dataset_offset, b0 = 4u * per;
        Normals6Scaled z;
        float r[6];
        if (VEC) {
            // the period's four blocks in registers, then three sectors: plane j = pairs (2j, 2j+1) of the four blocks
            float q[4][6];
#pragma unroll
            for (int ph = 0; ph < 4; ph++) {
                philox_pairs_lg2(b0 + (uint32_t)ph, tg, dg, STREAM_STEP, a.key, rk, z);
                euler6_rec(x, n, p, c0, h, z, a.max_steps, r);
#pragma unroll
                for (int i = 0; i < 6; i++) q[ph][i] = r[i];
            }
            if (rec && DDM_CHECK(a.stats, ds < a.n_datasets && trial < a.n_trials &&
                                              6u * a.rec_g * (b0 / a.rec_g) + 4u * a.rec_g + 2u * (b0 % a.rec_g) + 8u <= a.rec_stride)) {
                const uint32_t g = a.rec_g, it = b0 / g, sl0 = b0 - it * g;  // chunk, first of the four lanes' slots
                float *dst = a.rec_path + ((uint64_t)ds * a.n_trials + trial) * a.rec_stride + 6u * g * it + 2u * sl0;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    float4 *d = reinterpret_cast<float4 *>(dst + 2u * g * j);
                    d[0] = make_float4(q[0][2 * j], q[0][2 * j + 1], q[1][2 * j], q[1][2 * j + 1]);
                    d[1] = make_float4(q[2][2 * j], q[2][2 * j + 1], q[3][2 * j], q[3][2 * j + 1]);
                }
            }
        } else {
            float *dst = a.rec_path + ((uint64_t)ds * a.n_trials + trial) * a.rec_stride + n0;
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
// production, second kernel: G lanes per trial finish the recorded path, the row in registers
// --------------------------------------------------------------------------------------------
// A trial's n_obs noise normals come six per Philox block of its aux stream, so the natural owner of observations
// 6b .. 6b+5 is the lane that draws block b.  G lanes share a trial (32 / G trials per warp at a time), lane sl takes
// blocks sl, sl + G, ... (at most POST_ITER of them) and keeps its <= 6 * POST_ITER observations in registers from the
// load of the recorded states to the store of the standardised row: 24 contiguous bytes per lane and block in
// (G lanes: 24 G contiguous bytes, whole sectors) and the same out, mean / variance by shuffle reduction.  Round 1's
// kernel staged the row in shared memory so that every access ran k = sl, sl + G, ... -- two shared-memory round trips
// per observation and four loops over the row: 354 warp instructions per trial, instruction-bound at 75 % issue with
// the memory system half idle.
// In: the recording kernel's chunked layout (see there) makes a lane's j-th pair of a block 8 contiguous bytes per
// lane over the G lanes.  Out: rows are in time order, so a chunk's 6 G values take one trip through a per-warp
// shared-memory buffer (three STS.64, three LDS.64 per lane) and leave as 8 (16 for float64) contiguous bytes per
// lane -- straight 24-byte-stride stores touched every sector three times (measured: 81 instead of 25 sectors per row).
//   PAIRS: n_obs is even -- rows are 8-byte aligned, observations move two at a time.
//   CHUNKED: the recorded rows are in the chunked layout (n_obs a multiple of 8); otherwise in time order.
constexpr int POST_ITER = 5;
#ifndef DDM_POST_MIN_BLOCKS
#define DDM_POST_MIN_BLOCKS 4  // 64 registers; measured 6 % ahead of 3 blocks at 72
#endif

template <bool OUT64, int G, bool PAIRS, bool CHUNKED>
__global__ void __launch_bounds__(256, DDM_POST_MIN_BLOCKS) evidence_post_kernel(const EvidenceArgs a, uint64_t total) {
    constexpr unsigned TPW = 32u / G;  // trials per warp
    __shared__ float tbuf_all[8][192];  // per warp: one chunk of every trial the warp works on, 6 floats per lane
    float *tbuf = tbuf_all[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u, sl = lane & (G - 1u), sub = lane / G;
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
        const float *row_in = a.rec_path + g * a.rec_stride;
        // 1. recorded states -> evidence units, held at the final evidence after the crossing, + sigma1 * z_noise[k]
        //    (positions past the trial's last step may never have been written by the recording kernel: selected
        //    away, never used in arithmetic that survives)
        float v[POST_ITER][6];
        float sum = 0.f;
#pragma unroll
        for (int it = 0; it < POST_ITER; it++) {
            const uint32_t b = sl + (uint32_t)it * G, k0 = 6u * b;
#pragma unroll
            for (int i = 0; i < 6; i++) v[it][i] = 0.f;
            if (b < n_blocks) {
                float rec[6];
                if (CHUNKED) {  // padded rows: every block of every chunk lies inside the row's allocation
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const float2 t = __ldg(reinterpret_cast<const float2 *>(row_in + 6u * G * it + 2u * G * j + 2u * sl));
                        rec[2 * j] = t.x;
                        rec[2 * j + 1] = t.y;
                    }
                } else if (PAIRS) {
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        float2 t = make_float2(0.f, 0.f);
                        if (k0 + 2u * j < a.n_obs) t = __ldg(reinterpret_cast<const float2 *>(row_in + k0) + j);
                        rec[2 * j] = t.x;
                        rec[2 * j + 1] = t.y;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 6; i++) rec[i] = (k0 + i < a.n_obs) ? __ldg(row_in + k0 + i) : 0.f;
                }
                float z[6];
                philox_normals6_f32(b, trial + a.trial_offset, ds + a.dataset_offset, STREAM_AUX, a.key, z);
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const uint32_t k = k0 + i;
                    if (k < a.n_obs) {
                        const float e = (k < nj) ? __fmul_rn(__fadd_rn(rec[i], h), u) : evj;
                        v[it][i] = __fmaf_rn(sigma1, z[i], e);
                        sum += v[it][i];
                    }
                }
            }
        }
        // 2. mean (and variance) over the row
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, o);
        const float mean = sum * inv_n;
        float scale = 1.f, shift = 0.f;
        if (a.mode == 1) {
            float ssd = 0.f;
#pragma unroll
            for (int it = 0; it < POST_ITER; it++) {
                const uint32_t k0 = 6u * (sl + (uint32_t)it * G);
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const float d = v[it][i] - mean;
                    if (k0 + i < a.n_obs) ssd = __fmaf_rn(d, d, ssd);
                }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) ssd += __shfl_xor_sync(FULL_MASK, ssd, o);
            scale = 1.f / sqrtf(ssd * inv_n);
            shift = mean;
        }
        // 3. the row
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
        if (PAIRS) {
            // a chunk at a time through the warp's buffer: lane (sub, sl) puts its block at words 6 lane .. 6 lane + 5 and
            // takes pair sl of each plane m: observations 6 G it + 2 G m + 2 sl (+1)
#pragma unroll
            for (int it = 0; it < POST_ITER; it++) {
                if (6u * G * it >= a.n_obs) break;  // warp-uniform
#pragma unroll
                for (int j = 0; j < 3; j++)
                    reinterpret_cast<float2 *>(tbuf + 6u * lane)[j] =
                        make_float2((v[it][2 * j] - shift) * scale, (v[it][2 * j + 1] - shift) * scale);
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 3; m++) {
                    const uint32_t k = 6u * G * it + 2u * G * m + 2u * sl;
                    const float2 o = *reinterpret_cast<const float2 *>(tbuf + 6u * G * sub + 2u * G * m + 2u * sl);
                    if (valid && k < a.n_obs) {
                        const uint64_t at = row + 2u + k;
                        if (OUT64) *reinterpret_cast<double2 *>(reinterpret_cast<double *>(a.out) + at) = make_double2((double)o.x, (double)o.y);
                        else *reinterpret_cast<float2 *>(reinterpret_cast<float *>(a.out) + at) = o;
                    }
                }
                __syncwarp();
            }
        } else if (valid) {
#pragma unroll
            for (int it = 0; it < POST_ITER; it++) {
                const uint32_t k0 = 6u * (sl + (uint32_t)it * G);
                const uint64_t at = row + 2u + k0;
                {
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        if (k0 + i < a.n_obs) {
                            const float o0 = (v[it][i] - shift) * scale;
                            if (OUT64) reinterpret_cast<double *>(a.out)[at + i] = (double)o0;
                            else reinterpret_cast<float *>(a.out)[at + i] = o0;
                        }
                    }
                }
            }
        }
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
    const uint64_t cap = (uint64_t)sm_count * DDM_POST_MIN_BLOCKS * 8;
    if (grid > cap) grid = cap;
    const bool pairs = (a.n_obs & 1u) == 0u, chunked = (a.n_obs & 7u) == 0u;
    if (chunked && a.rec_g != (uint32_t)G) return cudaErrorInvalidValue;  // the recording kernel wrote for another split
    if (out64) {
        if (chunked) evidence_post_kernel<true, G, true, true><<<(unsigned)grid, 256, 0, s>>>(a, total);
        else if (pairs) evidence_post_kernel<true, G, true, false><<<(unsigned)grid, 256, 0, s>>>(a, total);
        else evidence_post_kernel<true, G, false, false><<<(unsigned)grid, 256, 0, s>>>(a, total);
    } else {
        if (chunked) evidence_post_kernel<false, G, true, true><<<(unsigned)grid, 256, 0, s>>>(a, total);
        else if (pairs) evidence_post_kernel<false, G, true, false><<<(unsigned)grid, 256, 0, s>>>(a, total);
        else evidence_post_kernel<false, G, false, false><<<(unsigned)grid, 256, 0, s>>>(a, total);
    }
    return cudaGetLastError();
}

// Lanes per trial in the post kernel, and with it the chunk of the recorded rows' layout (6 G observations): the
// widest split that keeps a lane's share of the row within POST_ITER Philox blocks.
uint32_t evidence_lanes_per_trial(uint32_t n_obs) {
    const uint32_t n_blocks = (n_obs + 5u) / 6u;
    return n_blocks <= 8u * POST_ITER ? 8u : (n_blocks <= 16u * POST_ITER ? 16u : 32u);
}

// Floats per recorded row: whole chunks in the chunked layout (n_obs a multiple of 8), n_obs otherwise.
uint32_t evidence_rec_stride(uint32_t n_obs) {
    if (n_obs & 7u) return n_obs;
    const uint32_t g = evidence_lanes_per_trial(n_obs), chunk = 6u * g;
    return (n_obs + chunk - 1u) / chunk * chunk;
}

cudaError_t launch_evidence_post(const EvidenceArgs &a, bool out64, uint64_t total, int sm_count, cudaStream_t s) {
    if (total == 0) return cudaSuccess;
    // the widest split that keeps a lane's share of the row within POST_ITER Philox blocks: 8 lanes per trial up to
    // 240 observations (the reference's 200), 16 up to 480 (its 400), the whole warp up to 960
    const uint32_t n_blocks = (a.n_obs + 5u) / 6u;
    if (n_blocks <= 8u * POST_ITER) return launch_evidence_post_g<8>(a, out64, total, sm_count, s);
    if (n_blocks <= 16u * POST_ITER) return launch_evidence_post_g<16>(a, out64, total, sm_count, s);
    if (n_blocks <= 32u * POST_ITER) return launch_evidence_post_g<32>(a, out64, total, sm_count, s);
    return cudaErrorInvalidValue;
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
