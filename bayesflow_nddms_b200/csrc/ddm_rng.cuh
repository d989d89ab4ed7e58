// ddm_rng.cuh -- counter-based normals in registers (sm_100a).
//
// Philox4x32-10 (Salmon et al., SC'11).  key = (seed_lo, seed_hi) is uniform for
// a launch, so the ten round keys live in uniform registers / the constant bank
// and a round costs 2 IMAD.WIDE.U32 + 2 LOP3 per lane.
// counter = (block, trial, dataset, stream):
//   stream 0 ("step"):  block b yields the normals of Euler steps 4b..4b+3
//   stream 1 ("aux"):   normal 0 = the ext-data normal z_ext (drawn after the
//                       loop in the reference, single_trial_alpha_not_scaled.py:131),
//                       normal 1+i = i-th candidate of the redraw-until-positive
//                       boundary / dc loop (:113-116, :932-935)
// so step normals sit at fixed counters no matter how many pre-draws a trial needs.
//
// u32 -> normal: Box-Muller on 23-bit uniforms built by bit injection (no I2F):
//   f  = as_float((w & 0x7fffff) | 0x3f800000)            in [1,2)
//   u  = f - (1 - 2^-24) = (2m+1)/2^24                     in (0,1), exact
//   t  = f' - 1.5                                          in [-.5,.5) revolutions
//   z_even = sqrt(-2 ln u) cos(2 pi t),  z_odd = sqrt(-2 ln u) sin(2 pi t)
// with MUFU lg2 / sqrt / sin / cos.  |z| <= 5.77 (P(|Z|>5.77) = 8e-9).
// oracle/ddm_oracle.c:orc_philox_normals4 is the fp64 value of the same map.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ddm {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

constexpr uint32_t STREAM_STEP = 0u;
constexpr uint32_t STREAM_AUX = 1u;

struct PhiloxKey {
    uint32_t k0, k1;
};

template <int ROUNDS = 10>
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t (&o)[4]) {
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;  // IMAD.WIDE.U32
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;  // LOP3
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += PHILOX_W0;  // uniform datapath: key is launch-uniform
        k1 += PHILOX_W1;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__device__ __forceinline__ float mufu_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sin(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_cos(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float bits_to_unit12(uint32_t w) {
    return __uint_as_float((w & 0x007fffffu) | 0x3f800000u);  // [1,2)
}

// One Box-Muller pair with the radius pre-scaled:  s = sqrt(|k * lg2(u)|) where the
// caller folds its noise scale into k = -2 ln2 * (sqrt(dt)*dc)^2, so an Euler step is
// x = fma(s, trig, x + c0) with no separate multiply.  k = -2 ln2 gives unit normals.
// |.| guards MUFU.LG2's absolute error near u -> 1 (a slightly positive lg2 would
// otherwise make the sqrt argument negative).
__device__ __forceinline__ void box_muller_scaled(uint32_t wa, uint32_t wb, float k, float &s,
                                                  float &c, float &sn) {
    const float u = __fadd_rn(bits_to_unit12(wa), -0.99999994f);  // (2m+1)/2^24, exact
    const float l = mufu_lg2(u);
    s = mufu_sqrt(fabsf(__fmul_rn(k, l)));
    const float a = __fmaf_rn(bits_to_unit12(wb), 6.2831853071795865f, -9.4247779607693797f);  // 2pi*(f-1.5)
    c = mufu_cos(a);
    sn = mufu_sin(a);
}

constexpr float NEG_2LN2 = -1.3862943611198906f;

// The four unit normals of Philox block (block, trial, dataset, stream), fp32 production map.
__device__ __forceinline__ void philox_normals4_f32(uint32_t block, uint32_t trial, uint32_t dataset,
                                                    uint32_t stream, PhiloxKey key, float (&z)[4]) {
    uint32_t w[4];
    philox4x32<10>(block, trial, dataset, stream, key.k0, key.k1, w);
    float s, c, sn;
    box_muller_scaled(w[0], w[1], NEG_2LN2, s, c, sn);
    z[0] = __fmul_rn(s, c);
    z[1] = __fmul_rn(s, sn);
    box_muller_scaled(w[2], w[3], NEG_2LN2, s, c, sn);
    z[2] = __fmul_rn(s, c);
    z[3] = __fmul_rn(s, sn);
}

// fp64 validation map: same bits, libdevice log/sincospi in double.
__device__ __forceinline__ void philox_normals4_f64(uint32_t block, uint32_t trial, uint32_t dataset,
                                                    uint32_t stream, PhiloxKey key, double (&z)[4]) {
    uint32_t w[4];
    philox4x32<10>(block, trial, dataset, stream, key.k0, key.k1, w);
#pragma unroll
    for (int p = 0; p < 2; p++) {
        const double u = (2.0 * (double)(w[2 * p] & 0x7fffffu) + 1.0) / 16777216.0;
        const double t = (double)(w[2 * p + 1] & 0x7fffffu) / 8388608.0 - 0.5;
        const double r = sqrt(-2.0 * log(u));
        double sn, c;
        sincospi(2.0 * t, &sn, &c);
        z[2 * p] = r * c;
        z[2 * p + 1] = r * sn;
    }
}

}  // namespace ddm
