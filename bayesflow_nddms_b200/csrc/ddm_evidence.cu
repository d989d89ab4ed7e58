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
//   production (fp32)         two kernels.  (1) The persistent refill kernel of ddm_kernels.cu in its RECORD form
//                             steps the trials (same lanes, same Philox counters as DDM_MODEL_BASIC) and stores the
//                             first n_obs states of each trial, six per block as three 8-byte stores, plus step
//                             count, final state and the (rt, choice) pair.  (2) evidence_post_kernel: a warp per
//                             trial turns the recorded states into the observed path -- evidence units, held at the
//                             final evidence after the crossing, noise normals from the aux Philox stream (six per
//                             lane), mean / variance by shuffle reduction -- and writes the row with coalesced
//                             stores.  800 B/trial of output make this the one DDM path where stores matter.
//   evidence_generic_kernel   validation (fp64): one thread per trial, the reference's operation order and
//                             left-to-right sums (numba's array_mean / array_var); shared-increment mode.
//   dataset_stats / finalize  mode 2's second pass, and the dtype conversion of the validation path.
#include "ddm_kernels.cuh"

namespace ddm {


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
        const uint32_t nj = (uint32_t)a.steps[g];
        const float h = a.dconst[ds].v[2], u = a.dconst[ds].v[3];
        const float sigma1 = (float)a.params[(size_t)ds * 6 + 5];
        const float evj = __fmul_rn(__fadd_rn(a.rec_xfinal[g], h), u);
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
            const double2 pr = a.pairs[g];  // (rt, choice) from the stepping kernel, reference fp64 arithmetic
            if (OUT64) {
                double *o = reinterpret_cast<double *>(a.out) + row;
                o[0] = pr.x;
                o[1] = pr.y;
            } else {
                float *o = reinterpret_cast<float *>(a.out) + row;
                o[0] = (float)pr.x;
                o[1] = (float)pr.y;
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
