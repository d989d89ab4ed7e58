// ddm_rng.cuh -- counter-based normals in registers (sm_100a).
//
// Philox4x32-10 (Salmon et al., SC'11).  key = (seed_lo, seed_hi) is uniform for
// a launch, so the ten round keys live in uniform registers / the constant bank
// and a round costs 2 IMAD.WIDE.U32 + 2 LOP3 per lane.
// counter = (stream | dataset [bits 32..55] << 8, block, trial, dataset [low 32 bits]); one block yields SIX normals:
//   stream 0 ("step"):  block b yields the normals of Euler steps 6b..6b+5
//   stream 1 ("aux"):   normal 0 = the ext-data normal z_ext (drawn after the
//                       loop in the reference, single_trial_alpha_not_scaled.py:131),
//                       normal 1+i = i-th candidate of the redraw-until-positive
//                       boundary / dc loop (:113-116, :932-935)
// so step normals sit at fixed counters no matter how many pre-draws a trial needs.
// The word order is chosen for the stepping loop: the block index -- the only word that changes while a trial
// runs -- sits in word 1, which enters round 0 through an XOR, not through a multiply.  Of the first three rounds'
// six products, four then depend on (stream, trial, dataset) alone: the compiler's loop-invariant code motion forms
// them once per entry of the stepping loop (checked in SASS: 16 IMAD.WIDE per block of the tile kernel's loop; with
// the block index in word 0 -- rounds 1 and 2 -- only two can leave the loop, 18 per block).  Same function, same
// ten rounds: only which word counts what.
//
// 128 bits -> three Box-Muller pairs of 21-bit uniforms (126 bits used, none twice).
// IMAD.WIDE issues at one per 4 clocks per scheduler on B200 (measured, bench.py
// --microbench), so Philox blocks are the scarce resource; 21-bit fields cost a few
// LOP3/SHF more than 23-bit ones and buy 6 instead of 4 normals per block.
//   pair 0: U = w0[2..22]  T = w1[2..22]        (fields already at mantissa position)
//   pair 1: U = w2[2..22]  T = w3[2..22]
//   pair 2: U = w0[23..31,0..1] ++ w1[23..31,0]  T = the same bits of w2, w3
//           (mantissa bits 2..12 = rotl(wa,11), bits 13..22 = rotl(wb,22))
// u32 field m (21 bits) -> by mantissa injection, no I2F:
//   f  = as_float(0x3f800000 | m << 2)                 in [1,2)
//   u  = f - (1 - 2^-22) = (2m+1)/2^22                 in (0,1), exact
//   t  = f' - 1.5                                      in [-.5,.5) revolutions
//   z_even = sqrt(-2 ln u) cos(2 pi t),  z_odd = sqrt(-2 ln u) sin(2 pi t)
// with MUFU lg2 / sqrt / sin / cos.  |z| <= 5.53 (P(|Z|>5.53) = 3e-8).
// oracle/ddm_oracle.c:orc_philox_normals6 is the fp64 value of the same map.
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

// The launch-uniform key with its ten round keys precomputed on the host: they sit in the
// kernel-parameter constant bank, so a round's key injection is an operand of the LOP3, not
// an extra uniform-datapath add per round per block.
struct PhiloxKey {
    uint32_t rk[20];  // rk[2r] = k0 + r*W0, rk[2r+1] = k1 + r*W1
    // Bits 8..31 of counter word 0 (bits 0..7 = stream): the high part of the 64-bit global dataset
    // index, launch-uniform (a launch never straddles a multiple of 2^32 datasets, ddm_capi.cu:
    // build_args), so long runs roll over into fresh counters instead of running out of them.
    uint32_t c3_hi;
};

__host__ __device__ inline PhiloxKey make_philox_key(uint32_t k0, uint32_t k1, uint32_t index_hi = 0u) {
    PhiloxKey k;
    for (int r = 0; r < 10; r++) {
        k.rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u;
        k.rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
    k.c3_hi = index_hi << 8;
    return k;
}

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

// Rounds of the simulator's generator.  10 is Random123's and cuRAND's default and what every shipped number is
// measured with; the build option -DDDM_PHILOX_ROUNDS=7 (Philox4x32-7: the fewest rounds that pass BigCrush in
// Salmon et al., SC'11, table 2) exists to measure what the round count costs (bench.py: philox7_variant) -- the
// stepping loop is bound by the 20 IMAD.WIDE of a 10-round block.  The known-answer tests (philox4x32<10>) and the
// CPU oracle are 10-round: a 7-round library is checked by the distribution tests only.
#ifndef DDM_PHILOX_ROUNDS
#define DDM_PHILOX_ROUNDS 10
#endif
static_assert(DDM_PHILOX_ROUNDS >= 7 && DDM_PHILOX_ROUNDS <= 10, "Philox4x32 rounds: 7 .. 10");

// The simulator's counter layout (see the header): block (block, trial, dataset, stream) of a launch's key.
__device__ __forceinline__ void philox4x32_rk(uint32_t block, uint32_t trial, uint32_t dataset, uint32_t stream,
                                              const PhiloxKey &key, uint32_t (&o)[4]) {
    uint32_t c0 = stream | key.c3_hi, c1 = block, c2 = trial, c3 = dataset;
#pragma unroll
    for (int r = 0; r < DDM_PHILOX_ROUNDS; r++) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        // (word ^ key) first: where the word is loop-invariant in a caller (see the header) so is this term
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ (c1 ^ key.rk[2 * r]);
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ (c3 ^ key.rk[2 * r + 1]);
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
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

constexpr int NORMALS_PER_BLOCK = 6;
constexpr uint32_t FIELD_MASK = 0x007ffffcu;  // mantissa bits 2..22
constexpr uint32_t ONE_BITS = 0x3f800000u;

// (x & FIELD_MASK) | ONE_BITS in one LOP3: float in [1,2) on a 2^-21 grid
__device__ __forceinline__ float field_to_unit12(uint32_t x) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "r"(FIELD_MASK), "r"(ONE_BITS));  // (a & b) | c
    return __uint_as_float(r);
}

// the leftover bits of two words as a third field: bits 2..12 from rotl(wa,11), 13..22 from rotl(wb,22)
__device__ __forceinline__ uint32_t leftover_field(uint32_t wa, uint32_t wb) {
    const uint32_t ra = __funnelshift_l(wa, wa, 11), rb = __funnelshift_l(wb, wb, 22);
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(0x00001ffcu), "r"(ra), "r"(rb));  // a ? b : c (bit select)
    return r;
}

// One Box-Muller pair in "lg2 units": s = sqrt(|lg2 u|), so the pair (s*c, s*sn) is N(0, 1/(2 ln 2))
// -- the kernels fold sqrt(2 ln 2) and their own noise scale into the (per-dataset or per-trial)
// constants of the state instead of multiplying every radius: the simulator state is
// x = (evidence - bound/2) / (sqrt(dt) * dc * sqrt(2 ln 2)), and an Euler step is
// x += fma(s, trig, c0) with no scale multiply at all.  |.| is a free operand modifier; it also
// guards MUFU.LG2's absolute error near u -> 1 (a slightly positive lg2).
__device__ __forceinline__ void box_muller_lg2(uint32_t fu, uint32_t ft, float &s, float &c, float &sn) {
    const float u = __fadd_rn(field_to_unit12(fu), -0.99999976158142089844f);  // (2m+1)/2^22, exact
    s = mufu_sqrt(fabsf(mufu_lg2(u)));
    const float a = __fmaf_rn(field_to_unit12(ft), 6.2831853071795865f, -9.4247779607693797f);  // 2pi*(f-1.5)
    c = mufu_cos(a);
    sn = mufu_sin(a);
}

// The same map with its two register constants handed in.  LOP3 and FFMA take one immediate each, so the exponent
// bits and 2 pi sit in registers, and ptxas re-creates them (two IMAD.MOV on the FMA-heavy pipe, the loop's
// bottleneck) in every block of a stepping loop unless they come out of a volatile asm (pinned_rng_consts).
struct RngConsts {
    uint32_t one_bits;
    float two_pi;
};
__device__ __forceinline__ RngConsts pinned_rng_consts() {
    // ptxas folds a plain mov of an immediate; a value derived from a special register it cannot
    unsigned eq, id;
    asm volatile("mov.u32 %0, %%lanemask_eq;" : "=r"(eq));
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(id));
    const uint32_t one = eq >> id;  // = 1
    RngConsts k;
    k.one_bits = 0x3f7fffffu + one;
    k.two_pi = __uint_as_float(0x40c90fdau + one);
    return k;
}
__device__ __forceinline__ float field_to_unit12(uint32_t x, const RngConsts &k) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "r"(FIELD_MASK), "r"(k.one_bits));  // (a & b) | c
    return __uint_as_float(r);
}
__device__ __forceinline__ void box_muller_lg2(uint32_t fu, uint32_t ft, const RngConsts &k, float &s, float &c, float &sn) {
    const float u = __fadd_rn(field_to_unit12(fu, k), -0.99999976158142089844f);
    s = mufu_sqrt(fabsf(mufu_lg2(u)));
    const float a = __fmaf_rn(field_to_unit12(ft, k), k.two_pi, -9.4247779607693797f);
    c = mufu_cos(a);
    sn = mufu_sin(a);
}

constexpr float SQRT_2LN2 = 1.1774100225154747f;         // unit normal = SQRT_2LN2 * s * trig
constexpr double SQRT_2LN2_D = 1.17741002251547469101;

// Radii and trig factors of the three pairs of a Philox block (the simulator's inner block).
struct Normals6Scaled {
    float s[3], c[3], sn[3];
};

__device__ __forceinline__ void philox_pairs_lg2(uint32_t block, uint32_t trial, uint32_t dataset,
                                                 uint32_t stream, const PhiloxKey &key, Normals6Scaled &o) {
    uint32_t w[4];
    philox4x32_rk(block, trial, dataset, stream, key, w);
    box_muller_lg2(w[0], w[1], o.s[0], o.c[0], o.sn[0]);
    box_muller_lg2(w[2], w[3], o.s[1], o.c[1], o.sn[1]);
    box_muller_lg2(leftover_field(w[0], w[1]), leftover_field(w[2], w[3]), o.s[2], o.c[2], o.sn[2]);
}

__device__ __forceinline__ void philox_pairs_lg2(uint32_t block, uint32_t trial, uint32_t dataset, uint32_t stream,
                                                 const PhiloxKey &key, const RngConsts &k, Normals6Scaled &o) {
    uint32_t w[4];
    philox4x32_rk(block, trial, dataset, stream, key, w);
    box_muller_lg2(w[0], w[1], k, o.s[0], o.c[0], o.sn[0]);
    box_muller_lg2(w[2], w[3], k, o.s[1], o.c[1], o.sn[1]);
    box_muller_lg2(leftover_field(w[0], w[1]), leftover_field(w[2], w[3]), k, o.s[2], o.c[2], o.sn[2]);
}

// The six unit normals of Philox block (block, trial, dataset, stream), fp32 production map.
__device__ __forceinline__ void philox_normals6_f32(uint32_t block, uint32_t trial, uint32_t dataset,
                                                    uint32_t stream, const PhiloxKey &key, float (&z)[6]) {
    Normals6Scaled o;
    philox_pairs_lg2(block, trial, dataset, stream, key, o);
#pragma unroll
    for (int p = 0; p < 3; p++) {
        const float r = __fmul_rn(SQRT_2LN2, o.s[p]);
        z[2 * p] = __fmul_rn(r, o.c[p]);
        z[2 * p + 1] = __fmul_rn(r, o.sn[p]);
    }
}

// fp64 validation map: same bits, libdevice log/sincospi in double.
__device__ __forceinline__ void philox_normals6_f64(uint32_t block, uint32_t trial, uint32_t dataset,
                                                    uint32_t stream, const PhiloxKey &key, double (&z)[6]) {
    uint32_t w[4];
    philox4x32_rk(block, trial, dataset, stream, key, w);
    const uint32_t fu[3] = {w[0], w[2], leftover_field(w[0], w[1])};
    const uint32_t ft[3] = {w[1], w[3], leftover_field(w[2], w[3])};
#pragma unroll
    for (int p = 0; p < 3; p++) {
        const double u = (2.0 * (double)((fu[p] & FIELD_MASK) >> 2) + 1.0) / 4194304.0;
        const double t = (double)((ft[p] & FIELD_MASK) >> 2) / 2097152.0 - 0.5;
        const double r = sqrt(-2.0 * log(u));
        double sn, c;
        sincospi(2.0 * t, &sn, &c);
        z[2 * p] = r * c;
        z[2 * p + 1] = r * sn;
    }
}

}  // namespace ddm
