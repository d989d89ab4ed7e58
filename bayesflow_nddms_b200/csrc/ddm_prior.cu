// ddm_prior.cu -- device-side batched prior sampler.
//
// Replaces B calls of draw_prior() (basic_ddm_dc.py:62-80, single_trial_alpha_not_scaled.py:78-102,
// :899-923 (_alt), :1205-1232 (_scale), retired_models/basic_ddm_eta_dc.py:54-75,
// retired_models/basic_ddm_dc_evidence.py:61-82) -- 0.84-1.0 ms each in the reference (three or four
// scipy truncnorm.rvs calls) -- by one launch that writes the (B, P) float64 parameter matrix straight
// into the simulator's parameter arena, in the reference's column order.
//
// Same distributions: Normal by Box-Muller, truncated normals by inverse CDF
// (mean + sd * normcdfinv(Fa + u (Fb - Fa))), Beta(2,2) as the median of three uniforms (the 2nd order
// statistic of 3 U(0,1) is Beta(2,2) exactly), Uniform by scaling.  fp64 throughout.
// Randomness: Philox4x32-10, counter = (column block, draw index lo, draw index hi, stream 2); block j
// gives the two 53-bit uniforms of column j, block 8 the third uniform of the Beta draw.
#include "ddm_kernels.cuh"

namespace ddm {

constexpr uint32_t STREAM_PRIOR = 2u;

__device__ __forceinline__ void prior_uniforms(uint32_t block, uint64_t draw, const PhiloxKey &key, double &ua, double &ub) {
    uint32_t w[4];
    philox4x32_rk(block, (uint32_t)draw, (uint32_t)(draw >> 32), STREAM_PRIOR, key, w);
    // 53-bit uniforms in (0,1): ((hi27 * 2^26 + lo26) + 0.5) / 2^53
    ua = (((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) + 0.5) / 9007199254740992.0;
    ub = (((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) + 0.5) / 9007199254740992.0;
}

__device__ __forceinline__ double prior_normal(uint32_t block, uint64_t draw, const PhiloxKey &key, double mean, double sd) {
    double ua, ub;
    prior_uniforms(block, draw, key, ua, ub);
    double sn, c;
    sincospi(2.0 * ub, &sn, &c);
    return mean + sd * (sqrt(-2.0 * log(ua)) * c);
}

__device__ __forceinline__ double prior_truncnorm(uint32_t block, uint64_t draw, const PhiloxKey &key, double mean,
                                                  double sd, double low, double upp) {
    double ua, ub;
    prior_uniforms(block, draw, key, ua, ub);
    const double fa = normcdf((low - mean) / sd), fb = normcdf((upp - mean) / sd);
    const double x = mean + sd * normcdfinv(fa + ua * (fb - fa));
    return fmin(fmax(x, low), upp);
}

__device__ __forceinline__ double prior_uniform(uint32_t block, uint64_t draw, const PhiloxKey &key, double lo, double hi) {
    double ua, ub;
    prior_uniforms(block, draw, key, ua, ub);
    return lo + (hi - lo) * ua;
}

__device__ __forceinline__ double prior_beta22(uint64_t draw, const PhiloxKey &key) {
    double u0, u1, u2, unused;
    prior_uniforms(2u, draw, key, u0, u1);
    prior_uniforms(8u, draw, key, u2, unused);
    return fmax(fmin(u0, u1), fmin(fmax(u0, u1), u2));  // median of three
}

// prior ids (include/ddm_b200.h: enum ddm_prior)
__global__ void prior_kernel(double *__restrict__ params, int prior, uint32_t n_params, uint64_t n_draws,
                             uint64_t draw_offset, const PhiloxKey key) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_draws) return;
    const uint64_t d = draw_offset + i;
    double *p = params + i * n_params;
    p[0] = prior_normal(0u, d, key, 0.0, 2.0);                  // drift | mu_drift ~ N(0, 2)
    p[1] = prior_truncnorm(1u, d, key, 1.0, 0.5, 0.0, 10.0);    // alpha | mu_alpha ~ TN(1, .5; 0, 10)
    p[2] = prior_beta22(d, key);                                // beta ~ Beta(2, 2)
    p[3] = (prior == 7) ? 0.0 : prior_truncnorm(3u, d, key, 0.5, 0.25, 0.0, 1.5);  // ter ~ TN(.5, .25; 0, 1.5); sweep: 0
    switch (prior) {
    case 0: case 7:  // basic, sweep: dc
        p[4] = prior_truncnorm(4u, d, key, 1.0, 0.5, 0.0, 10.0);
        break;
    case 8:          // evidence: dc, sigma1 ~ U(0, 5)
        p[4] = prior_truncnorm(4u, d, key, 1.0, 0.5, 0.0, 10.0);
        p[5] = prior_uniform(5u, d, key, 0.0, 5.0);
        break;
    case 6:          // eta: eta ~ TN(1, .5; 0, 3), dc
        p[4] = prior_truncnorm(4u, d, key, 1.0, 0.5, 0.0, 3.0);
        p[5] = prior_truncnorm(5u, d, key, 1.0, 0.5, 0.0, 10.0);
        break;
    default:         // alpha family: std_alpha|std_dc ~ TN(1, .5; 0, 3), dc|mu_dc, sigma1 ~ U(0, 5)(, gamma ~ U(0, 2))
        p[4] = prior_truncnorm(4u, d, key, 1.0, 0.5, 0.0, 3.0);
        p[5] = prior_truncnorm(5u, d, key, 1.0, 0.5, 0.0, 10.0);
        p[6] = prior_uniform(6u, d, key, 0.0, 5.0);
        if (prior == 3) p[7] = prior_uniform(7u, d, key, 0.0, 2.0);
        break;
    }
}

cudaError_t launch_prior(double *params, int prior, uint32_t n_params, uint64_t n_draws, uint64_t draw_offset,
                         const PhiloxKey &key, cudaStream_t s) {
    if (n_draws == 0) return cudaSuccess;
    prior_kernel<<<(unsigned)((n_draws + 127) / 128), 128, 0, s>>>(params, prior, n_params, n_draws, draw_offset, key);
    return cudaGetLastError();
}

}  // namespace ddm
