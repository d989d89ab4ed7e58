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

}  // namespace ddm
