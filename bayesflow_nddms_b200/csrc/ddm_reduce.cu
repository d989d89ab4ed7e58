// On-device reduction of a resident batch: response-time histograms by choice (SURVEY section 8d, config C5:
// "outputs reduced on device (RT histogram + choice counts)").  The histogram is additive over datasets, so
// shards, chunks and GPUs can be checked against one another without moving the rows.
#include "ddm_kernels.cuh"

namespace ddm {

// rows: `cols` values per trial; column 0/1 = (rt, choice) [basic layout] or (signed rt, .) [signed layout].
// hist: [0, n_bins) upper-boundary responses, [n_bins, 2 n_bins) lower, then {missing / timed out, rt >= rt_max}.
template <typename T>
__global__ void __launch_bounds__(256) rt_histogram_kernel(const T *rows, uint64_t n_rows, uint32_t cols, bool basic_layout,
                                                           uint32_t n_bins, double inv_width, unsigned long long *hist) {
    extern __shared__ unsigned int h_smem[];
    const uint32_t n_cells = 2u * n_bins + 2u;
    for (uint32_t i = threadIdx.x; i < n_cells; i += blockDim.x) h_smem[i] = 0u;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_rows; g += stride) {
        const T *r = rows + g * cols;
        double rt;
        int choice;
        if (basic_layout) {
            rt = (double)r[0];
            const double c = (double)r[1];
            choice = (c > 0.0) - (c < 0.0);
        } else {
            const double v = (double)r[0];
            rt = fabs(v);
            choice = (v > 0.0) - (v < 0.0);
        }
        uint32_t cell;
        if (choice == 0) {
            cell = 2u * n_bins;
        } else {
            const double b = rt * inv_width;
            cell = (b >= (double)n_bins) ? 2u * n_bins + 1u : (uint32_t)b + (choice < 0 ? n_bins : 0u);
        }
        atomicAdd(&h_smem[cell], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_cells; i += blockDim.x)
        if (h_smem[i]) atomicAdd(&hist[i], (unsigned long long)h_smem[i]);
}

cudaError_t launch_rt_histogram(const void *rows, bool rows64, uint64_t n_rows, uint32_t cols, bool basic_layout, uint32_t n_bins,
                                double rt_max, unsigned long long *hist, int sm_count, cudaStream_t s) {
    if (n_rows == 0) return cudaSuccess;
    const size_t smem = (size_t)(2u * n_bins + 2u) * sizeof(unsigned int);
    uint64_t grid = (n_rows + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count * 8;  // a block's shared histogram costs one flush: keep blocks few and long-lived
    if (grid > cap) grid = cap;
    const double inv_width = (double)n_bins / rt_max;
    if (rows64)
        rt_histogram_kernel<double><<<(unsigned)grid, 256, smem, s>>>((const double *)rows, n_rows, cols, basic_layout, n_bins, inv_width, hist);
    else
        rt_histogram_kernel<float><<<(unsigned)grid, 256, smem, s>>>((const float *)rows, n_rows, cols, basic_layout, n_bins, inv_width, hist);
    return cudaGetLastError();
}

// Quantifies the production normal generator: draws Philox blocks (block, trial = thread, dataset,
// stream 0) exactly as the stepping kernels do, turns each into six normals with the production fp32 map
// (philox_normals6_f32: 21-bit fields, MUFU lg2/sqrt/sin/cos) and reduces them on the device to
//   hist[0 .. nb_abs)            counts of |z| in [k, k+1) * z_max / nb_abs        hist[nb_abs] = |z| >= z_max
//   hist[nb_abs+1 .. +nb_ang)    counts of the pair's angle atan2(z_odd, z_even) in nb_ang equal sectors
//   moments: sum z, sum z^2, sum z^3, sum z^4 (double)
__global__ void __launch_bounds__(256) normals_histogram_kernel(PhiloxKey key, uint64_t n_blocks, uint32_t blocks_per_thread,
                                                                uint32_t nb_abs, float inv_width, uint32_t nb_ang,
                                                                unsigned long long *hist, double *moments) {
    extern __shared__ unsigned int nh_smem[];
    const uint32_t n_cells = nb_abs + 1u + nb_ang;
    for (uint32_t i = threadIdx.x; i < n_cells; i += blockDim.x) nh_smem[i] = 0u;
    __syncthreads();
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double m1 = 0.0, m2 = 0.0, m3 = 0.0, m4 = 0.0;
    const uint64_t first = tid * blocks_per_thread;
    for (uint32_t b = 0; b < blocks_per_thread; b++) {  // block-uniform trip count: the flush below has barriers
        if (first + b < n_blocks) {
        float z[6];
        // counters laid out like a trial's: block b of (trial, dataset) = the low and high words of the thread index
        philox_normals6_f32(b, (uint32_t)tid, (uint32_t)(tid >> 32), STREAM_STEP, key, z);
        float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const float a = fabsf(z[k]);
            const float q = a * inv_width;
            const uint32_t cell = (q >= (float)nb_abs) ? nb_abs : (uint32_t)q;
            atomicAdd(&nh_smem[cell], 1u);
            const float zz = z[k] * z[k];
            s1 += z[k]; s2 += zz; s3 += zz * z[k]; s4 += zz * zz;
        }
#pragma unroll
        for (int pr = 0; pr < 3; pr++) {
            float ang = atan2f(z[2 * pr + 1], z[2 * pr]) * 0.15915494309189535f;  // revolutions in [-.5, .5]
            ang = (ang < 0.f ? ang + 1.f : ang) + 2.384185791015625e-07f;  // half a lattice step: lattice angles sit inside sectors
            uint32_t cell = (uint32_t)(ang * (float)nb_ang);
            if (cell >= nb_ang) cell = nb_ang - 1u;
            atomicAdd(&nh_smem[nb_abs + 1u + cell], 1u);
        }
        m1 += (double)s1; m2 += (double)s2; m3 += (double)s3; m4 += (double)s4;
        }
        if ((b & 1023u) == 1023u) {  // 32-bit shared counters: flush well before they can wrap
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < n_cells; i += blockDim.x) {
                const unsigned int v = nh_smem[i];
                if (v) { atomicAdd(&hist[i], (unsigned long long)v); nh_smem[i] = 0u; }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_cells; i += blockDim.x)
        if (nh_smem[i]) atomicAdd(&hist[i], (unsigned long long)nh_smem[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m1 += __shfl_xor_sync(FULL_MASK, m1, o);
        m2 += __shfl_xor_sync(FULL_MASK, m2, o);
        m3 += __shfl_xor_sync(FULL_MASK, m3, o);
        m4 += __shfl_xor_sync(FULL_MASK, m4, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&moments[0], m1); atomicAdd(&moments[1], m2); atomicAdd(&moments[2], m3); atomicAdd(&moments[3], m4);
    }
}

cudaError_t launch_normals_histogram(const PhiloxKey &key, uint64_t n_blocks, uint32_t nb_abs, double z_max, uint32_t nb_ang,
                                     unsigned long long *hist, double *moments, int sm_count, cudaStream_t s) {
    if (n_blocks == 0) return cudaSuccess;
    const uint32_t blocks_per_thread = 2048;  // every thread runs the loop in lock-step with its block (the flush barriers)
    const uint64_t threads = (n_blocks + blocks_per_thread - 1) / blocks_per_thread;
    const uint64_t grid = (threads + 255) / 256;
    (void)sm_count;
    const size_t smem = (size_t)(nb_abs + 1u + nb_ang) * sizeof(unsigned int);
    normals_histogram_kernel<<<(unsigned)grid, 256, smem, s>>>(key, n_blocks, blocks_per_thread, nb_abs, (float)((double)nb_abs / z_max),
                                                                nb_ang, hist, moments);
    return cudaGetLastError();
}

}  // namespace ddm
