// ddm_kernels.cuh -- shared device-side definitions of the DDM simulator kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ddm_rng.cuh"
#include "ddm_wire.cuh"

namespace ddm {

// Kernel families.  The model enum of include/ddm_b200.h maps onto these.
enum Kind : int {
    KIND_FIXED = 0,     // basic_ddm_dc: all trial constants are per-dataset
    KIND_BOUND = 1,     // per-trial boundary redraw (single_trial_alpha*, _scale, _scale2)
    KIND_DC = 2,        // per-trial diffusion coefficient redraw (_alt)
    KIND_TRIALWISE = 3, // per-trial supplied boundary + gathered group params (Stahl)
    KIND_DRIFT = 4,     // per-trial drift ~ N(mu_drift, eta) (basic_ddm_eta_dc)
    KIND_GENERAL = 5    // per-trial drift, boundary and dc + two external channels, three output columns
};

// Per-dataset constants in fp32, prepared once per upload by prep_kernel (fp64 math,
// rounded once).  The simulator state is measured in units of the step's noise scale,
// U = sqrt(dt)*dc*sqrt(2 ln 2) (see ddm_rng.cuh: box_muller_lg2).  v[] meaning per kind:
//   FIXED: c0=drift*dt/U, x0=bound*(beta-.5)/U, h=(bound/2)/U, U
//   BOUND: c0=drift*dt/U, (beta-.5)/U, bmu, bsd | .5/U, ext_sd, ext_gain, U
//   DC:    drift*dt/U1, bound*(beta-.5)/U1, (bound/2)/U1, mu_dc | std_dc, ext_sd, -, U1   (U1 = U at dc = 1)
//   DRIFT: mu_drift*dt/U, x0, h, U | eta*dt/U
struct __align__(16) DsConst {
    float v[8];
};

// KIND_GENERAL: per-dataset constants (fp64 math in prep_general_kernel, rounded once).  unit1 = sqrt(dt)*sqrt(2 ln 2)
// is the state unit at dc = 1; the per-trial unit is unit1 * dc_t.
//   0 drift_mu*dt/unit1  1 drift_sd*dt/unit1  2 bound_mu  3 bound_sd  4 dc_mu  5 dc_sd  6 (beta-.5)/unit1  7 .5/unit1
//   8 drift_mu  9 drift_sd  10 unit1
//   11..15 channel 1: A = -shift/scale, coefficients of drift_t, bound_t, dc_t over scale, sigma/scale
//   16..20 channel 2 likewise   21 output style
struct __align__(16) GenConst {
    float v[24];
};

constexpr uint32_t TILE_MAX_STEPS = (1u << 30) - 1u;  // tile kernel: a trial's result is (steps << 2 | choice + 1) in 32 bits
constexpr int REJECT_CAP_BLOCKS = 1024;  // ~6000 candidates before giving up on a redraw loop

struct RunArgs {
    // inputs
    const DsConst *dconst;   // [n_datasets]
    const GenConst *gconst;  // [n_datasets], KIND_GENERAL only
    const double *params;    // raw fp64 parameters [n_datasets * n_params] (group params for trialwise)
    const int32_t *group;    // trialwise: [n_trials]
    const double *bound_in;  // trialwise: [n_trials]
    const double *dbg_z;     // shared-increment mode
    const int64_t *dbg_off;
    uint64_t dbg_n;          // normals in dbg_z (reads past the end return 0 and are counted)
    // outputs
    void *out;               // float2 / double2 per trial
    int32_t *steps_out;      // optional
    float *rec_path;         // evidence models: [trial][rec_stride] centred, unit-scaled state after each step
    uint32_t rec_stride;     // floats per recorded row; rec_g: lanes per trial of the post kernel (chunk = 6 rec_g
    uint32_t rec_g;          // observations, see record_kernel)
    uint32_t n_obs;
    unsigned long long *work_counter;
    unsigned long long *stats;  // see StatSlot
    // shape
    uint64_t n_items;        // persistent: number of (dataset, tile) work items
    uint32_t n_datasets, n_trials, n_params;
    uint32_t tiles_per_dataset, tile;
    uint32_t dataset_offset;  // global index of dataset 0 (Philox counter word 3)
    uint32_t trial_offset;
    PhiloxKey key;
    uint32_t max_steps;
    int model;               // ddm_model
    int flags;
    int refill_threshold;
    double dt, sqrt_dt;
};

constexpr unsigned FULL_MASK = 0xffffffffu;

// compute-sanitizer is closed on the GPU pool this library is developed on, so the kernels carry their own bounds
// checks: a build with -DDDM_CHECKED (bayesflow_nddms_b200/_build.py: build_checked) verifies every shared-memory slot
// index, every output row index and every gathered input index before use and counts violations in the
// STAT_DBG_OVERRUN slot (ddm_stats.debug_overruns); a violating access is skipped.  scripts/r02_checked_run.py drives
// every kernel family through it.  In the shipped build the macro compiles to nothing.
#ifdef DDM_CHECKED
#define DDM_CHECK(stats, cond) ((cond) ? true : (atomicAdd((stats) + STAT_DBG_OVERRUN, 1ull), false))
#else
#define DDM_CHECK(stats, cond) true
#endif
// internal run flag (not part of the C ABI): the fp32 generic kernel uses the reference's formulas in
// float instead of the unit-scaled production arithmetic (a dataset with dc == 0 has no noise unit)
constexpr int FLAG_REFERENCE_ARITHMETIC = 1 << 30;
// internal: the persistent kernel writes ddm_wire.cuh records (4 or 8 bytes per trial) instead of output rows
constexpr int FLAG_WIRE_COMPACT = 1 << 29;

enum StatSlot { STAT_STEPS = 0, STAT_TIMEOUTS = 1, STAT_UPPER = 2, STAT_REJECT_CAP = 3, STAT_DBG_OVERRUN = 4, STAT_COUNT = 5 };

// ---- fp32 trial state shared by the persistent and the generic kernels ------------
struct TrialF32 {
    float x;    // (evidence - bound/2) / U  (centred state: one |x| < h compare per step)
    float h;    // (bound/2) / U
    float c0;   // drift*dt / U
    float u;    // U = sqrt(dt)*dc*sqrt(2 ln 2): evidence = (x + h) * U (validation output only)
    float ext;  // second output column (ext-data / boundary), decided at setup
    float ext2; // third output column (KIND_GENERAL)
};

template <int KIND>
__device__ __forceinline__ void trial_setup_f32(const DsConst &dc, uint32_t trial, uint32_t ds_global,
                                                const PhiloxKey &key, TrialF32 &t, uint32_t &cap_hits) {
    t.c0 = dc.v[0];
    if (KIND == KIND_DRIFT) {  // drift_trial = mu_drift + eta*z, one pre-draw (aux normal 1), no rejection
        float z[6];
        philox_normals6_f32(0u, trial, ds_global, STREAM_AUX, key, z);
        t.c0 = __fmaf_rn(dc.v[4], z[1], dc.v[0]);
        t.x = dc.v[1];
        t.h = dc.v[2];
        t.u = dc.v[3];
        t.ext = 0.f;
        return;
    }
    if (KIND == KIND_FIXED) {
        t.x = dc.v[1];
        t.h = dc.v[2];
        t.u = dc.v[3];
        t.ext = 0.f;
        return;
    }
    const float mu = (KIND == KIND_BOUND) ? dc.v[2] : dc.v[3];
    const float sd = (KIND == KIND_BOUND) ? dc.v[3] : dc.v[4];
    float z[6];
    philox_normals6_f32(0u, trial, ds_global, STREAM_AUX, key, z);
    const float z_ext = z[0];
    float latent = __fmaf_rn(sd, z[1], mu);
#pragma unroll
    for (int i = 2; i < 6; i++)
        if (!(latent > 0.f)) latent = __fmaf_rn(sd, z[i], mu);
    for (uint32_t j = 1; !(latent > 0.f); j++) {
        if (j >= REJECT_CAP_BLOCKS) {
            cap_hits++;
            latent = 1e-30f;
            break;
        }
        philox_normals6_f32(j, trial, ds_global, STREAM_AUX, key, z);
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (!(latent > 0.f)) latent = __fmaf_rn(sd, z[i], mu);
    }
    if (KIND == KIND_BOUND) {
        t.h = __fmul_rn(dc.v[4], latent);
        t.x = __fmul_rn(latent, dc.v[1]);
        t.u = dc.v[7];
        t.ext = __fmaf_rn(dc.v[5], z_ext, __fmul_rn(dc.v[6], latent));
    } else {  // KIND_DC: the per-trial diffusion coefficient rescales the whole state
        const float inv = __frcp_rn(latent);
        t.c0 = __fmul_rn(dc.v[0], inv);
        t.x = __fmul_rn(dc.v[1], inv);
        t.h = __fmul_rn(dc.v[2], inv);
        t.u = __fmul_rn(dc.v[7], latent);
        t.ext = __fmaf_rn(dc.v[5], z_ext, latent);
    }
}

// KIND_TRIALWISE (imputation_from_stahl_not_scaled.py:120-148): the boundary is supplied, the constants are the
// trial's participant's (BOUND layout: c0, (beta-.5)/U, -, -, .5/U, -, -, U).  bound == 0 gives h = x = 0: no step
// is taken and x >= h reports the upper boundary, as the reference does (ev >= bound_trial -> +ter).
__device__ __forceinline__ void trial_setup_trialwise(const DsConst &gc, float bound, TrialF32 &t) {
    t.c0 = gc.v[0];
    t.h = __fmul_rn(gc.v[4], bound);
    t.x = __fmul_rn(bound, gc.v[1]);
    t.u = gc.v[7];
    t.ext = bound;
    t.ext2 = 0.f;
}

// KIND_GENERAL: aux stream normal 0 = z_ext1, 1 = z_ext2, 2 = z_drift, boundary candidate i = 4 + 2i,
// dc candidate i = 5 + 2i (so block 0 carries the first candidate of each redraw loop).
__device__ __forceinline__ void trial_setup_general(const GenConst &g, uint32_t trial, uint32_t ds_global,
                                                    const PhiloxKey &key, TrialF32 &t, uint32_t &cap_hits) {
    float z[6];
    philox_normals6_f32(0u, trial, ds_global, STREAM_AUX, key, z);
    const float z1 = z[0], z2 = z[1], zd = z[2];
    const bool bound_fixed = g.v[3] == 0.f, dc_fixed = g.v[5] == 0.f;
    float bound_t = bound_fixed ? g.v[2] : __fmaf_rn(g.v[3], z[4], g.v[2]);
    float dc_t = dc_fixed ? g.v[4] : __fmaf_rn(g.v[5], z[5], g.v[4]);
    bool need_b = !bound_fixed && !(bound_t > 0.f), need_c = !dc_fixed && !(dc_t > 0.f);
    for (uint32_t j = 1; need_b || need_c; j++) {
        if (j >= REJECT_CAP_BLOCKS) {
            cap_hits++;
            if (need_b) bound_t = 1e-30f;
            if (need_c) dc_t = 1e-30f;
            break;
        }
        philox_normals6_f32(j, trial, ds_global, STREAM_AUX, key, z);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (need_b) { bound_t = __fmaf_rn(g.v[3], z[2 * i], g.v[2]); need_b = !(bound_t > 0.f); }
            if (need_c) { dc_t = __fmaf_rn(g.v[5], z[2 * i + 1], g.v[4]); need_c = !(dc_t > 0.f); }
        }
    }
    const float inv = __frcp_rn(dc_t);
    t.c0 = __fmul_rn(__fmaf_rn(g.v[1], zd, g.v[0]), inv);
    t.x = __fmul_rn(__fmul_rn(bound_t, g.v[6]), inv);
    t.h = __fmul_rn(__fmul_rn(bound_t, g.v[7]), inv);
    t.u = __fmul_rn(g.v[10], dc_t);
    const float drift_t = __fmaf_rn(g.v[9], zd, g.v[8]);
    t.ext = __fmaf_rn(g.v[15], z1, __fmaf_rn(g.v[14], dc_t, __fmaf_rn(g.v[13], bound_t, __fmaf_rn(g.v[12], drift_t, g.v[11]))));
    t.ext2 = __fmaf_rn(g.v[20], z2, __fmaf_rn(g.v[19], dc_t, __fmaf_rn(g.v[18], bound_t, __fmaf_rn(g.v[17], drift_t, g.v[16]))));
}

// Six predicated Euler steps from one Philox block.  `p` (0/1) = "still inside the boundaries and
// below max_steps"; a lane whose p is 0 is frozen (x, n keep their crossing values).  Written in
// PTX so that every step is exactly  FFMA, @p FADD, @p IADD, FSETP.AND  (no branches: the warp
// executes the block while any lane is alive, so skipping buys nothing).  The increment is formed
// first, inc = fma(s, trig, c0), then x += inc: the reference's own association ev + (t1 + t3)
// (basic_ddm_dc.py:98), one rounding at ulp(x) per step, and the x-chain is a single FADD.  (Adding
// c0 to x first would round the tiny drift term to x's grid the same way every step -- a systematic
// drift error of up to ulp(x)/2 per step.)  TAIL adds the
// n < max_steps test per step; the callers use it only for a trial's last, partial block.
#define DDM_STEP_G(S, T, C0, H)                \
    "fma.rn.f32 inc, " S ", " T ", " C0 ";\n\t" \
    "@q add.rn.f32 %0, %0, inc;\n\t"           \
    "@q add.u32 %1, %1, 1;\n\t"                \
    "abs.f32 ax, %0;\n\t"                      \
    "setp.lt.and.f32 q, ax, " H ", q;\n\t"
#define DDM_STEP(S, T) DDM_STEP_G(S, T, "%3", "%4")
#define DDM_STEP_TAIL(S, T) DDM_STEP(S, T) "setp.lt.and.u32 q, %1, %14, q;\n\t"

// No warp-level primitive in here: the one-thread-per-trial kernel calls this from divergent code.
// (A float step counter, to move the increments from the ALU pipe to the FMA pipes, measured 4 %
// slower on B200: the FMA-heavy pipe is as busy as the ALU pipe because of IMAD.WIDE.)
template <bool TAIL>
__device__ __forceinline__ void euler6(float &x, uint32_t &n, uint32_t &p, float c0, float h,
                                       const Normals6Scaled &z, uint32_t max_steps) {
    if (TAIL) {
        asm volatile("{\n\t.reg .pred q;\n\t.reg .f32 ax, inc;\n\t"
                     "setp.ne.u32 q, %2, 0;\n\t"
                     "setp.lt.and.u32 q, %1, %14, q;\n\t"
                     DDM_STEP_TAIL("%5", "%6") DDM_STEP_TAIL("%5", "%7") DDM_STEP_TAIL("%8", "%9")
                     DDM_STEP_TAIL("%8", "%10") DDM_STEP_TAIL("%11", "%12") DDM_STEP_TAIL("%11", "%13")
                     "selp.u32 %2, 1, 0, q;\n\t}"
                     : "+f"(x), "+r"(n), "+r"(p)
                     : "f"(c0), "f"(h), "f"(z.s[0]), "f"(z.c[0]), "f"(z.sn[0]), "f"(z.s[1]), "f"(z.c[1]), "f"(z.sn[1]),
                       "f"(z.s[2]), "f"(z.c[2]), "f"(z.sn[2]), "r"(max_steps));
    } else {
        asm volatile("{\n\t.reg .pred q;\n\t.reg .f32 ax, inc;\n\t"
                     "setp.ne.u32 q, %2, 0;\n\t"
                     DDM_STEP("%5", "%6") DDM_STEP("%5", "%7") DDM_STEP("%8", "%9")
                     DDM_STEP("%8", "%10") DDM_STEP("%11", "%12") DDM_STEP("%11", "%13")
                     "selp.u32 %2, 1, 0, q;\n\t}"
                     : "+f"(x), "+r"(n), "+r"(p)
                     : "f"(c0), "f"(h), "f"(z.s[0]), "f"(z.c[0]), "f"(z.sn[0]), "f"(z.s[1]), "f"(z.c[1]), "f"(z.sn[1]),
                       "f"(z.s[2]), "f"(z.c[2]), "f"(z.sn[2]), "r"(max_steps));
    }
}

// The persistent kernel's form of the same six steps: lane state "still stepping" travels as the
// warp's ballot `alive` (bit = lane), so the block needs no predicate<->register conversions:
// LOP3 (alive & lanemask -> predicate), the steps, ISETP (n < max_steps, folded into the predicate),
// VOTE.  Convergent code only.
__device__ __forceinline__ void euler6_warp(float &x, uint32_t &n, unsigned &alive, float c0, float h,
                                            const Normals6Scaled &z, uint32_t max_steps) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .f32 ax, inc;\n\t.reg .b32 t;\n\t"
                 "mov.u32 t, %%lanemask_eq;\n\t"
                 "and.b32 t, t, %2;\n\t"
                 "setp.ne.u32 q, t, 0;\n\t"
                 DDM_STEP("%5", "%6") DDM_STEP("%5", "%7") DDM_STEP("%8", "%9")
                 DDM_STEP("%8", "%10") DDM_STEP("%11", "%12") DDM_STEP("%11", "%13")
                 "setp.lt.and.u32 q, %1, %14, q;\n\t"
                 "vote.sync.ballot.b32 %2, q, 0xffffffff;\n\t}"
                 : "+f"(x), "+r"(n), "+r"(alive)
                 : "f"(c0), "f"(h), "f"(z.s[0]), "f"(z.c[0]), "f"(z.sn[0]), "f"(z.s[1]), "f"(z.c[1]), "f"(z.sn[1]),
                   "f"(z.s[2]), "f"(z.c[2]), "f"(z.sn[2]), "r"(max_steps));
}

// The same with the lane's own bit handed in (the tile kernel holds it in a register anyway; reading
// %lanemask_eq costs an S2R per block when the compiler does not hoist it).
__device__ __forceinline__ void euler6_warp_lb(float &x, uint32_t &n, unsigned &alive, unsigned lane_bit, float c0, float h,
                                               const Normals6Scaled &z, uint32_t max_steps) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .f32 ax, inc;\n\t.reg .b32 t;\n\t"
                 "and.b32 t, %15, %2;\n\t"
                 "setp.ne.u32 q, t, 0;\n\t"
                 DDM_STEP("%5", "%6") DDM_STEP("%5", "%7") DDM_STEP("%8", "%9")
                 DDM_STEP("%8", "%10") DDM_STEP("%11", "%12") DDM_STEP("%11", "%13")
                 "setp.lt.and.u32 q, %1, %14, q;\n\t"
                 "vote.sync.ballot.b32 %2, q, 0xffffffff;\n\t}"
                 : "+f"(x), "+r"(n), "+r"(alive)
                 : "f"(c0), "f"(h), "f"(z.s[0]), "f"(z.c[0]), "f"(z.sn[0]), "f"(z.s[1]), "f"(z.c[1]), "f"(z.sn[1]),
                   "f"(z.s[2]), "f"(z.c[2]), "f"(z.sn[2]), "r"(max_steps), "r"(lane_bit));
}

// The same block for the evidence-path models (record_kernel, ddm_evidence.cu): "still stepping" is the lane's own
// 0/1 word, and the state after each of the six steps is handed back (r[k] = x after step k+1; frozen at the
// crossing value once the lane has stopped), which the caller stores as the observed path.
#define DDM_STEPR(S, T, R) DDM_STEP_G(S, T, "%9", "%10") "mov.f32 " R ", %0;\n\t"
__device__ __forceinline__ void euler6_rec(float &x, uint32_t &n, uint32_t &p, float c0, float h, const Normals6Scaled &z,
                                           uint32_t max_steps, float (&r)[6]) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .f32 ax, inc;\n\t"
                 "setp.ne.u32 q, %2, 0;\n\t"
                 DDM_STEPR("%11", "%12", "%3") DDM_STEPR("%11", "%13", "%4") DDM_STEPR("%14", "%15", "%5")
                 DDM_STEPR("%14", "%16", "%6") DDM_STEPR("%17", "%18", "%7") DDM_STEPR("%17", "%19", "%8")
                 "setp.lt.and.u32 q, %1, %20, q;\n\t"
                 "selp.u32 %2, 1, 0, q;\n\t}"
                 : "+f"(x), "+r"(n), "+r"(p), "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5])
                 : "f"(c0), "f"(h), "f"(z.s[0]), "f"(z.c[0]), "f"(z.sn[0]), "f"(z.s[1]), "f"(z.c[1]), "f"(z.sn[1]),
                   "f"(z.s[2]), "f"(z.c[2]), "f"(z.sn[2]), "r"(max_steps));
}

// One Philox block of a trial: block index `blk` = n / 6 for a lane that is still stepping.
template <bool TAIL>
__device__ __forceinline__ void step_block_f32(uint32_t blk, uint32_t trial, uint32_t ds_global,
                                               const PhiloxKey &key, const TrialF32 &t, float &x, uint32_t &n,
                                               uint32_t &p, uint32_t max_steps) {
    Normals6Scaled z;
    philox_pairs_lg2(blk, trial, ds_global, STREAM_STEP, key, z);
    euler6<TAIL>(x, n, p, t.c0, t.h, z, max_steps);
}

// Final outputs of a finished fp32 trial, computed in fp64 with the reference's operation
// order so that, given the same step count, col0 equals the reference's double exactly
// (basic_ddm_dc.py:103  rt = n_steps*dt + tau;  single_trial_alpha_not_scaled.py:127,133-138).
template <bool BASIC>
__device__ __forceinline__ void trial_outputs(int flags, int choice, uint32_t n, double dt, double tau, double ext,
                                              double &o0, double &o1) {
    const double rt = __dmul_rn((double)n, dt);
    if (BASIC) {  // DDM_MODEL_BASIC
        o0 = __dadd_rn(rt, tau);
        o1 = (choice == 0) ? ((flags & 1) ? 1.0 : 0.0) : (double)choice;
    } else {
        o0 = (choice > 0) ? __dadd_rn(tau, rt) : (choice < 0 ? __dsub_rn(-tau, rt) : 0.0);
        o1 = ext;
    }
}

// ---- evidence-path variants (ddm_evidence.cu) -------------------------------------------------------
struct EvidenceArgs {
    const double *params;    // [n_datasets * 6]: drift, boundary, beta, tau, dc, sigma1
    const double *dbg_z;     // shared-increment mode (validation kernel)
    const int64_t *dbg_off;
    uint64_t dbg_n;
    void *out;               // rows of 2 + n_obs values, float64 or float32
    double *scratch;         // validation path: fp64 rows
    double *path_means;      // mode 2: per-trial mean of the noisy path [n_datasets * n_trials]
    // production path: products of the stepping kernel (record_kernel)
    const float *rec_path;   // [trial][rec_stride] centred, unit-scaled states (layout: record_kernel)
    uint32_t rec_stride, rec_g;
    const uint2 *rec_meta;   // per trial: ((steps << 2) | (choice + 1), final state as fp32 bits)
    const DsConst *dconst;   // per-dataset constants (v[2] = h, v[3] = U)
    unsigned long long *work_counter;
    unsigned long long *stats;
    uint64_t n_items;
    uint32_t n_datasets, n_trials, n_obs, tiles_per_dataset;
    uint32_t dataset_offset, trial_offset, max_steps;
    int mode;                // 0 raw noisy path, 1 per-trial z-score, 2 dataset-level standardisation
    int flags;
    PhiloxKey key;
    double dt, sqrt_dt;
};

cudaError_t launch_exact_sampler(const double *params, double *out, unsigned long long *stats, uint32_t n_datasets,
                                 uint32_t n_trials, uint32_t dataset_offset, uint32_t trial_offset, const PhiloxKey &key,
                                 cudaStream_t s);
cudaError_t launch_rt_histogram(const void *rows, bool rows64, uint64_t n_rows, uint32_t cols, bool basic_layout, uint32_t n_bins,
                                double rt_max, unsigned long long *hist, int sm_count, cudaStream_t s);
cudaError_t launch_normals_histogram(const PhiloxKey &key, uint64_t n_blocks, uint32_t nb_abs, double z_max, uint32_t nb_ang,
                                     unsigned long long *hist, double *moments, int sm_count, cudaStream_t s);
cudaError_t launch_evidence_post(const EvidenceArgs &a, bool out64, uint64_t total, int sm_count, cudaStream_t s);
uint32_t evidence_lanes_per_trial(uint32_t n_obs);
uint32_t evidence_rec_stride(uint32_t n_obs);
cudaError_t launch_evidence_generic(const EvidenceArgs &a, bool buffer_src, uint64_t total, cudaStream_t s);
cudaError_t launch_evidence_dataset_stats(const double *path_means, double *ds_stats, uint32_t n_datasets,
                                          uint32_t n_trials, cudaStream_t s);
cudaError_t launch_evidence_finalize(const void *src, bool src64, void *dst, bool dst64, const double *ds_stats,
                                     uint64_t total, uint32_t cols, uint32_t n_trials, bool standardize, cudaStream_t s);

// device-side prior sampler (ddm_prior.cu)
cudaError_t launch_prior(double *params, int prior, uint32_t n_params, uint64_t n_draws, uint64_t draw_offset,
                         const PhiloxKey &key, cudaStream_t s);

// launchers (ddm_kernels.cu)
cudaError_t launch_prep(const double *params, DsConst *dconst, uint32_t n_datasets, uint32_t n_params,
                        int model, double dt, cudaStream_t s);
cudaError_t launch_persistent(const RunArgs &a, int kind, bool out64, int grid, int block, cudaStream_t s);
// the tile-staged persistent kernel (production): a.tile_cap and the shared-memory size come from tile_kernel_config
cudaError_t launch_tile(const RunArgs &a, int kind, bool out64, int grid, int block, size_t smem, cudaStream_t s);
size_t tile_kernel_smem_bytes(int kind, int block);
int tile_kernel_max_blocks_per_sm(int kind, bool out64, int block, size_t smem);
uint32_t tile_kernel_max_tile(int kind);
cudaError_t launch_record(const RunArgs &a, int grid, int block, cudaStream_t s);  // evidence models' stepping kernel (ddm_evidence.cu)
int record_max_blocks_per_sm(int block);
int record_block_size();
cudaError_t launch_prep_general(const double *params, GenConst *gconst, uint32_t n_datasets, double dt, cudaStream_t s);
// small launches (every kind but KIND_GENERAL): one thread per trial, speculative six-step blocks
cudaError_t launch_latency(const RunArgs &a, int kind, bool out64, uint64_t total_trials, cudaStream_t s);
cudaError_t launch_generic(const RunArgs &a, int kind, bool f64, bool buffer_src, bool out64,
                           uint64_t total_trials, cudaStream_t s);
cudaError_t launch_export_normals(PhiloxKey key, uint32_t dataset, uint32_t trial, uint32_t stream,
                                  uint32_t first, uint32_t count, bool f64, double *out_dev, cudaStream_t s);
cudaError_t launch_philox_blocks(const uint32_t *ctr, const uint32_t *key, uint32_t *out, int64_t n,
                                 cudaStream_t s);
int persistent_max_blocks_per_sm(int kind, bool out64, int block);
int persistent_block_size();

}  // namespace ddm
