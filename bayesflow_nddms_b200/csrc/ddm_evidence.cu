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
//   evidence_warp_kernel      production (fp32): a warp owns 32 trials; lanes step them in lock-step and
//                             record the first n_obs evidence values in shared memory ([trial][k], padded
//                             to an odd stride: conflict-free while recording); the warp then finishes each
//                             trial co-operatively: noise normals (aux Philox stream, six per lane), mean and
//                             variance by shuffle reduction, coalesced row stores.  800 B/trial of output make
//                             this the one DDM kernel where stores matter.
//   evidence_generic_kernel   validation (fp64): one thread per trial, the reference's operation order and
//                             left-to-right sums (numba's array_mean / array_var); shared-increment mode.
//   dataset_stats / finalize  mode 2's second pass, and the dtype conversion of the validation path.
#include "ddm_kernels.cuh"

namespace ddm {

constexpr int EV_MAX_BLOCKS_PER_LANE = 4;  // n_obs <= 32 * 6 * 4 = 768

// --------------------------------------------------------------------------------------------
// production: warp per 32 trials
// --------------------------------------------------------------------------------------------
template <bool OUT64>
__global__ void __launch_bounds__(128) evidence_warp_kernel(const EvidenceArgs a) {
    extern __shared__ float ev_smem[];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    const uint32_t stride = a.n_obs + 1u;  // odd or even, +1 breaks the power-of-two stride
    float *path = ev_smem + (size_t)warp * 32u * stride;
    const uint32_t cols = 2u + a.n_obs;
    const uint32_t n_blocks = (a.n_obs + 5u) / 6u;

    unsigned long long acc_steps = 0;
    uint32_t acc_timeouts = 0, acc_upper = 0;

    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1ull);
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w >= a.n_items) break;
        const uint32_t ds = (uint32_t)w / a.tiles_per_dataset;
        const uint32_t t0 = ((uint32_t)w - ds * a.tiles_per_dataset) * 32u;
        const uint32_t trial = t0 + lane;
        const bool valid = trial < a.n_trials;
        const uint32_t ds_g = ds + a.dataset_offset, trial_g = trial + a.trial_offset;

        const double *prm = a.params + (size_t)ds * 6;
        const double drift = prm[0], boundary = prm[1], beta = prm[2], tau = prm[3], dcoef = prm[4];
        const float sigma1 = (float)prm[5];
        TrialF32 t;
        const double U = a.sqrt_dt * dcoef * SQRT_2LN2_D;  // state unit (ddm_rng.cuh: box_muller_lg2)
        t.c0 = (float)(drift * a.dt / U);
        t.h = (float)(0.5 * boundary / U);
        t.u = (float)U;
        t.ext = 0.f;
        float x = (float)(boundary * (beta - 0.5) / U);
        uint32_t n = 0, blk = 0;
        uint32_t p = (valid && (fabsf(x) < t.h) && (a.max_steps > 0u)) ? 1u : 0u;

        // ---- phase 1: step and record while any live trial is inside the observation window ----
        while (__any_sync(FULL_MASK, p != 0u && n < a.n_obs)) {
            Normals6Scaled z;
            philox_pairs_lg2(blk, trial_g, ds_g, STREAM_STEP, a.key, z);
#pragma unroll
            for (int i = 0; i < 6; i++) {
                if (p) {
                    const float inc = __fmaf_rn(z.s[i >> 1], (i & 1) ? z.sn[i >> 1] : z.c[i >> 1], t.c0);
                    x = __fadd_rn(x, inc);
                    if (n < a.n_obs) path[lane * stride + n] = __fmul_rn(__fadd_rn(x, t.h), t.u);
                    n++;
                    p = ((fabsf(x) < t.h) && (n < a.max_steps)) ? 1u : 0u;
                }
            }
            blk++;
        }
        // ---- phase 2: the rest of the trial needs no recording ----
        while (__any_sync(FULL_MASK, p != 0u)) {
            step_block_f32<true>(blk, trial_g, ds_g, a.key, t, x, n, p, a.max_steps);
            blk++;
        }
        __syncwarp();
        const float ev_final = __fmul_rn(__fadd_rn(x, t.h), t.u);
        const int choice = (x >= t.h) ? 1 : ((x <= -t.h) ? -1 : 0);
        if (valid) {
            acc_steps += n;
            acc_timeouts += (choice == 0);
            acc_upper += (choice > 0);
        }

        // ---- phase 3: the warp finishes one trial at a time ----
        const uint32_t n_valid = min(32u, a.n_trials - t0);
        for (uint32_t j = 0; j < n_valid; j++) {
            const uint32_t nj = __shfl_sync(FULL_MASK, n, j);
            const float evj = __shfl_sync(FULL_MASK, ev_final, j);
            const int chj = __shfl_sync(FULL_MASK, choice, j);
            const uint32_t trial_j = t0 + j + a.trial_offset;
            float vals[EV_MAX_BLOCKS_PER_LANE * 6];
            float sum = 0.f;
#pragma unroll
            for (int bi = 0; bi < EV_MAX_BLOCKS_PER_LANE; bi++) {
                const uint32_t b = lane + 32u * bi;
                if (b < n_blocks) {
                    float z[6];
                    philox_normals6_f32(b, trial_j, ds_g, STREAM_AUX, a.key, z);
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        const uint32_t k = 6u * b + i;
                        float v = 0.f;
                        if (k < a.n_obs) {
                            const float base = (k < nj) ? path[j * stride + k] : evj;
                            v = __fmaf_rn(sigma1, z[i], base);
                            sum += v;
                        }
                        vals[bi * 6 + i] = v;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, o);
            const float mean = sum / (float)a.n_obs;
            float scale = 1.f, shift = 0.f;
            if (a.mode == 1) {
                float ssd = 0.f;
#pragma unroll
                for (int bi = 0; bi < EV_MAX_BLOCKS_PER_LANE; bi++) {
                    const uint32_t b = lane + 32u * bi;
#pragma unroll
                    for (int i = 0; i < 6; i++)
                        if (b < n_blocks && 6u * b + i < a.n_obs) {
                            const float d = vals[bi * 6 + i] - mean;
                            ssd = __fmaf_rn(d, d, ssd);
                        }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ssd += __shfl_xor_sync(FULL_MASK, ssd, o);
                scale = 1.f / sqrtf(ssd / (float)a.n_obs);
                shift = mean;
            }
            const uint64_t row = ((uint64_t)ds * a.n_trials + t0 + j) * cols;
            if (lane == 0) {
                const double rt = __dadd_rn(__dmul_rn((double)nj, a.dt), tau);
                if (OUT64) {
                    double *o = reinterpret_cast<double *>(a.out) + row;
                    o[0] = rt;
                    o[1] = (double)chj;
                } else {
                    float *o = reinterpret_cast<float *>(a.out) + row;
                    o[0] = (float)rt;
                    o[1] = (float)chj;
                }
                if (a.mode == 2) a.path_means[(uint64_t)ds * a.n_trials + t0 + j] = (double)mean;
            }
#pragma unroll
            for (int bi = 0; bi < EV_MAX_BLOCKS_PER_LANE; bi++) {
                const uint32_t b = lane + 32u * bi;
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const uint32_t k = 6u * b + i;
                    if (b < n_blocks && k < a.n_obs) {
                        const float v = (vals[bi * 6 + i] - shift) * scale;
                        if (OUT64) reinterpret_cast<double *>(a.out)[row + 2 + k] = (double)v;
                        else reinterpret_cast<float *>(a.out)[row + 2 + k] = v;
                    }
                }
            }
        }
        __syncwarp();
    }
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
__global__ void evidence_finalize_kernel(const Src *__restrict__ src, Dst *__restrict__ dst,
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
size_t evidence_smem_per_warp(uint32_t n_obs) { return (size_t)32 * (n_obs + 1) * sizeof(float); }

cudaError_t launch_evidence_warp(const EvidenceArgs &a, bool out64, int grid, int warps_per_block, cudaStream_t s) {
    const size_t smem = evidence_smem_per_warp(a.n_obs) * warps_per_block;
    cudaError_t e;
    if (out64) {
        if ((e = cudaFuncSetAttribute(evidence_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        evidence_warp_kernel<true><<<grid, 32 * warps_per_block, smem, s>>>(a);
    } else {
        if ((e = cudaFuncSetAttribute(evidence_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        evidence_warp_kernel<false><<<grid, 32 * warps_per_block, smem, s>>>(a);
    }
    return cudaGetLastError();
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
