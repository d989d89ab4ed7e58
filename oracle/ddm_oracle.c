/*
 * ddm_oracle.c -- CPU oracle for the DDM trial-simulator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in bayesflow_nddms_b200/ may call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load it.  It is a plain-C, fp64 restatement of the
 * reference's Euler-Maruyama trial loops, following the reference's operation
 * order exactly (no FMA contraction: build with -ffp-contract=off):
 *
 *   M0  basic_ddm_dc.py:85-112 (diffusion_trial), :114-125 (simulate_trials)
 *   M1  single_trial_alpha_not_scaled.py:107-142, :144-155
 *   ALT single_trial_alpha_not_scaled.py:926-961   (per-trial dc)
 *   SCALE :1237-1272 (ext = N(gamma*bound, sigma1)), SCALE2 :1471-1506 (gamma==2)
 *   FINE  :1710-1722 (M1 called with dt=.001, max_steps=4000)
 *   M2  imputation_from_stahl_not_scaled.py:120-148 (supplied per-trial bound)
 *   ETA retired_models/basic_ddm_eta_dc.py:80-120 (per-trial drift ~ N(mu_drift, eta))
 *   GENERAL  the retired zoo's two-latent / two-channel scripts in one parametrisation, e.g.
 *            retired_models/single_trial_drift_dc5.py:90-155, single_trial_alpha_dc.py:109-175
 *
 * Parity pin: the reference has no golden vectors, so the
 * oracle is pinned against outputs of the reference itself: tests/golden/
 * make_golden.py runs the reference's verbatim functions under numba with
 * np.random.seed(s) called inside jitted code; numba's stream is MT19937 +
 * the legacy polar method (numba cpython/randomimpl.py), which this file
 * restates (orc_mt_*), so oracle(seed) == reference(seed) bit for bit.
 *
 * The normal source is pluggable: an explicit buffer of standard normals
 * ("shared increments"), the MT19937/polar stream, or the Philox4x32-10
 * counter stream the CUDA kernels use (fp64 "ideal" transform of the same
 * bits).  Consumption order per trial:
 *   M0/M2: z_step[0..n-1];  M1/SCALE/SCALE2: z_bound[0..k], z_step, z_ext;
 *   ALT: z_dc[0..k], z_step, z_ext.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* model ids -- must match include/ddm_b200.h */
enum {
    ORC_MODEL_BASIC = 0,
    ORC_MODEL_ALPHA = 1,
    ORC_MODEL_ALPHA_DC = 2,
    ORC_MODEL_ALPHA_SCALE = 3,
    ORC_MODEL_ALPHA_SCALE2 = 4,
    ORC_MODEL_TRIALWISE = 5,
    ORC_MODEL_ETA = 6,
    ORC_MODEL_GENERAL = 7
};
/* flags -- must match include/ddm_b200.h */
enum { ORC_FLAG_TIMEOUT_CHOICE_ONE = 1 };

/* ------------------------------------------------------------------ */
/* MT19937 + legacy polar gauss: the stream numba's np.random.normal()  */
/* draws from after np.random.seed(s) inside jitted code.               */
/* ------------------------------------------------------------------ */
typedef struct {
    uint32_t mt[624];
    int idx;
    int has_gauss;
    double gauss;
} orc_mt;

ORC_EXPORT void orc_mt_seed(orc_mt *st, uint32_t seed) {
    st->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        st->mt[i] = 1812433253u * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (uint32_t)i;
    st->idx = 624;
    st->has_gauss = 0;
    st->gauss = 0.0;
}

static void mt_refill(orc_mt *st) {
    uint32_t *mt = st->mt;
    for (int k = 0; k < 624; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        uint32_t v = mt[(k + 397) % 624] ^ (y >> 1);
        if (y & 1u) v ^= 0x9908b0dfu;
        mt[k] = v;
    }
    st->idx = 0;
}

static inline uint32_t mt_u32(orc_mt *st) {
    if (st->idx >= 624) mt_refill(st);
    uint32_t y = st->mt[st->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

ORC_EXPORT double orc_mt_double(orc_mt *st) {
    uint32_t a = mt_u32(st) >> 5, b = mt_u32(st) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

ORC_EXPORT double orc_mt_gauss(orc_mt *st) {
    if (st->has_gauss) {
        st->has_gauss = 0;
        return st->gauss;
    }
    double x1, x2, r2;
    do {
        x1 = 2.0 * orc_mt_double(st) - 1.0;
        x2 = 2.0 * orc_mt_double(st) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    double f = sqrt(-2.0 * log(r2) / r2);
    st->gauss = f * x1;
    st->has_gauss = 1;
    return f * x2;
}

ORC_EXPORT void orc_mt_normals(uint32_t seed, double *out, size_t n) {
    orc_mt st;
    orc_mt_seed(&st, seed);
    for (size_t i = 0; i < n; i++) out[i] = orc_mt_gauss(&st);
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011), the counter stream of the CUDA   */
/* kernels.  key = (seed_lo, seed_hi); ctr = (stream, block, trial,     */
/* dataset).  Pinned by the Random123 known-answer vectors in tests/.   */
/* ------------------------------------------------------------------ */
ORC_EXPORT void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* fp64 "ideal" value of the 6 normals a Philox block encodes (three Box-Muller pairs of
 * 21-bit uniforms; field layout documented in bayesflow_nddms_b200/csrc/ddm_rng.cuh).  The
 * CUDA kernels compute the same map in fp32 with MUFU lg2/sqrt/sin/cos:
 *   u = (2*m + 1) / 2^22  in (0,1),  t = m' / 2^21 - 1/2
 *   z_even = sqrt(-2 ln u) cos(2 pi t),  z_odd = sqrt(-2 ln u) sin(2 pi t)   */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline uint32_t field21(uint32_t w) { return (w >> 2) & 0x1fffffu; } /* bits 2..22 */

static inline uint32_t leftover21(uint32_t wa, uint32_t wb) {
    /* mantissa bits 2..12 <- rotl(wa,11), bits 13..22 <- rotl(wb,22) */
    uint32_t ra = rotl32(wa, 11), rb = rotl32(wb, 22);
    return field21((ra & 0x00001ffcu) | (rb & ~0x00001ffcu));
}

ORC_EXPORT void orc_philox_normals6(uint64_t seed, uint32_t block, uint32_t trial,
                                    uint32_t dataset, uint32_t stream, double z[6]) {
    uint32_t ctr[4] = {stream, block, trial, dataset}; /* ddm_rng.cuh: the kernels' counter layout */
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t w[4];
    orc_philox4x32_10(ctr, key, w);
    const double two_pi = 6.283185307179586476925286766559;
    uint32_t mu[3] = {field21(w[0]), field21(w[2]), leftover21(w[0], w[1])};
    uint32_t mt[3] = {field21(w[1]), field21(w[3]), leftover21(w[2], w[3])};
    for (int p = 0; p < 3; p++) {
        double u = (2.0 * (double)mu[p] + 1.0) / 4194304.0;
        double t = (double)mt[p] / 2097152.0 - 0.5;
        double r = sqrt(-2.0 * log(u));
        z[2 * p] = r * cos(two_pi * t);
        z[2 * p + 1] = r * sin(two_pi * t);
    }
}

/* ------------------------------------------------------------------ */
/* Normal sources                                                       */
/* ------------------------------------------------------------------ */
enum { SRC_BUFFER = 0, SRC_MT = 1, SRC_PHILOX = 2 };
enum { PH_STREAM_STEP = 0, PH_STREAM_AUX = 1 };

typedef struct {
    int kind;
    /* buffer */
    const double *buf;
    size_t pos, len;
    int overrun;
    /* mt */
    orc_mt mt;
    /* philox: position is set per trial and per phase */
    uint64_t seed;
    uint32_t trial, dataset;
    uint32_t stream;    /* current stream */
    uint32_t next_idx;  /* index of next normal within the stream */
    double cache[6];
    uint32_t cache_block;
    int cache_valid;
} orc_src;

static inline double src_next(orc_src *s) {
    switch (s->kind) {
    case SRC_BUFFER:
        if (s->pos >= s->len) { s->overrun = 1; return 0.0; }
        return s->buf[s->pos++];
    case SRC_MT:
        return orc_mt_gauss(&s->mt);
    default: {
        uint32_t blk = s->next_idx / 6u;
        if (!s->cache_valid || s->cache_block != blk) {
            orc_philox_normals6(s->seed, blk, s->trial, s->dataset, s->stream, s->cache);
            s->cache_block = blk;
            s->cache_valid = 1;
        }
        uint32_t j = s->next_idx - 6u * blk;
        s->next_idx++;
        return s->cache[j];
    }
    }
}

static inline void src_seek(orc_src *s, uint32_t stream, uint32_t idx) {
    if (s->kind != SRC_PHILOX) return;
    s->stream = stream;
    s->next_idx = idx;
    s->cache_valid = 0;
}

/* ------------------------------------------------------------------ */
/* One trial.  Returns everything the parity tests look at.             */
/* ------------------------------------------------------------------ */
typedef struct {
    double out0, out1;   /* the two columns the reference stacks */
    double out2;         /* third column (general model: second external channel) */
    double evidence;     /* final evidence */
    double bound;        /* boundary used by this trial */
    int64_t n_steps;
    int choice;          /* +1 / -1 / 0 (timeout) */
    int64_t consumed;    /* normals consumed (buffer/MT sources) */
} orc_trial;

/* The Euler-Maruyama loop shared by every variant.
 * basic_ddm_dc.py:95-101: operation order t1=drift*dt; t2=sqrt(dt)*dc;
 * t3=t2*z; ev = ev + (t1 + t3); the step counter is a double. */
static inline void euler_loop(double drift, double bound, double beta, double dc,
                              double dt, double max_steps, orc_src *src,
                              double *ev_out, double *n_out) {
    double n_steps = 0.0;
    double evidence = bound * beta;
    while ((evidence > 0) && (evidence < bound) && (n_steps < max_steps)) {
        double z = src_next(src);
        double t1 = drift * dt;
        double t2 = sqrt(dt) * dc;
        double t3 = t2 * z;
        evidence = evidence + (t1 + t3);
        n_steps += 1.0;
    }
    *ev_out = evidence;
    *n_out = n_steps;
}

static void trial_run(int model, const double *p, double bound_in, double dt,
                      double max_steps, int flags, orc_src *src, orc_trial *o) {
    size_t pos0 = src->pos;
    double ev, n;
    memset(o, 0, sizeof(*o));
    if (model == ORC_MODEL_BASIC) {
        /* p = [drift, boundary, beta, tau, dc]  basic_ddm_dc.py:85-112 */
        src_seek(src, PH_STREAM_STEP, 0);
        euler_loop(p[0], p[1], p[2], p[4], dt, max_steps, src, &ev, &n);
        double rt = n * dt + p[3];
        int choice;
        if (ev >= p[1]) choice = 1;
        else if (ev <= 0) choice = -1;
        else choice = (flags & ORC_FLAG_TIMEOUT_CHOICE_ONE) ? 1 : 0; /* D8 */
        o->out0 = rt;
        o->out1 = (double)choice;
        o->choice = (ev >= p[1]) ? 1 : (ev <= 0 ? -1 : 0);
        o->bound = p[1];
    } else if (model == ORC_MODEL_GENERAL) {
        /* Canonical parameters (include/ddm_b200.h: DDM_MODEL_GENERAL):
         *  0 drift_mu 1 drift_sd | 2 bound_mu 3 bound_sd | 4 dc_mu 5 dc_sd | 6 beta 7 tau
         *  8..13  ext1: coefficient of drift_t, bound_t, dc_t; sigma; shift; scale
         *  14..19 ext2 likewise | 20 order of the pre-draws | 21 number of ext channels | 22 output style
         * A latent with sd == 0 is fixed and consumes no normal; drift is drawn once, boundary and dc are
         * redrawn until positive (single_trial_drift_dc5.py:99-105, single_trial_alpha_dc.py:115-125). */
        static const int ORDER[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
        int ord = (int)p[20];
        if (ord < 0 || ord > 5) ord = 0;
        double lat[3] = {p[0], p[2], p[4]};
        uint32_t cand[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++) {
            int which = ORDER[ord][k];
            double mu = p[2 * which], sd = p[2 * which + 1];
            if (sd == 0.0) continue;
            for (;;) {
                /* aux stream: drift z = normal 2; boundary candidate i = 4 + 2i; dc candidate i = 5 + 2i */
                uint32_t idx = (which == 0) ? 2u : (which == 1 ? 4u + 2u * cand[1] : 5u + 2u * cand[2]);
                src_seek(src, PH_STREAM_AUX, idx);
                lat[which] = mu + sd * src_next(src);
                cand[which]++;
                if (which == 0 || lat[which] > 0) break;
            }
        }
        double drift_t = lat[0], bound_t = lat[1], dc_t = lat[2];
        src_seek(src, PH_STREAM_STEP, 0);
        euler_loop(drift_t, bound_t, p[6], dc_t, dt, max_steps, src, &ev, &n);
        double rt = n * dt;
        int n_ext = (int)p[21];
        double ext[2] = {0.0, 0.0};
        for (int c = 0; c < 2 && c < n_ext; c++) {
            const double *e = p + 8 + 6 * c;
            src_seek(src, PH_STREAM_AUX, (uint32_t)c);
            double loc = (e[0] * drift_t + e[1] * bound_t) + e[2] * dc_t;
            double temp = loc + e[3] * src_next(src);
            ext[c] = (temp - e[4]) / e[5];
        }
        o->choice = (ev >= bound_t) ? 1 : (ev <= 0 ? -1 : 0);
        if ((int)p[22] == 0) { /* (rt, choice, ext1) */
            o->out0 = rt + p[7];
            o->out1 = (double)o->choice;
            o->out2 = ext[0];
        } else {               /* (signed choicert, ext1, ext2) */
            o->out0 = (o->choice > 0) ? p[7] + rt : (o->choice < 0 ? -p[7] - rt : 0.0);
            o->out1 = ext[0];
            o->out2 = ext[1];
        }
        o->bound = bound_t;
    } else if (model == ORC_MODEL_ETA) {
        /* p = [mu_drift, alpha, beta, ter, eta, dc]  retired_models/basic_ddm_eta_dc.py:80-107;
         * one pre-draw: drift_trial = mu_drift + eta*z (aux normal 1), then the basic loop */
        src_seek(src, PH_STREAM_AUX, 1);
        double drift_trial = p[0] + p[4] * src_next(src);
        src_seek(src, PH_STREAM_STEP, 0);
        euler_loop(drift_trial, p[1], p[2], p[5], dt, max_steps, src, &ev, &n);
        double rt = n * dt + p[3];
        int choice;
        if (ev >= p[1]) choice = 1;
        else if (ev <= 0) choice = -1;
        else choice = (flags & ORC_FLAG_TIMEOUT_CHOICE_ONE) ? 1 : 0; /* same unbound `choice` typo, :110-112 */
        o->out0 = rt;
        o->out1 = (double)choice;
        o->choice = (ev >= p[1]) ? 1 : (ev <= 0 ? -1 : 0);
        o->bound = p[1];
    } else if (model == ORC_MODEL_TRIALWISE) {
        /* p = [drift, beta, ter, dc]; imputation_from_stahl_not_scaled.py:120-148 */
        double bound = bound_in;
        src_seek(src, PH_STREAM_STEP, 0);
        euler_loop(p[0], bound, p[1], p[3], dt, max_steps, src, &ev, &n);
        double rt = n * dt;
        double choicert;
        if (ev >= bound) { choicert = p[2] + rt; o->choice = 1; }
        else if (ev <= 0) { choicert = -p[2] - rt; o->choice = -1; }
        else { choicert = 0; o->choice = 0; }
        o->out0 = choicert;
        o->out1 = bound;
        o->bound = bound;
    } else {
        /* 7/8-parameter family: single_trial_alpha_not_scaled.py:107-142 etc.
         * p = [drift, mu_alpha|alpha, beta, ter, std_alpha|std_dc, dc|mu_dc, sigma1(, gamma)] */
        double drift = p[0], beta = p[2], ter = p[3], sigma1 = p[6];
        double bound, dc, latent, gain;
        src_seek(src, PH_STREAM_AUX, 1); /* aux normal 0 is reserved for z_ext */
        if (model == ORC_MODEL_ALPHA_DC) {
            double dc_trial;
            for (;;) {
                dc_trial = p[5] + p[4] * src_next(src);
                if (dc_trial > 0) break;
            }
            bound = p[1]; dc = dc_trial; latent = dc_trial; gain = 1.0;
        } else {
            double bound_trial;
            for (;;) {
                bound_trial = p[1] + p[4] * src_next(src);
                if (bound_trial > 0) break;
            }
            bound = bound_trial; dc = p[5]; latent = bound_trial;
            gain = (model == ORC_MODEL_ALPHA_SCALE) ? p[7]
                 : (model == ORC_MODEL_ALPHA_SCALE2) ? 2.0 : 1.0;
        }
        src_seek(src, PH_STREAM_STEP, 0);
        euler_loop(drift, bound, beta, dc, dt, max_steps, src, &ev, &n);
        double rt = n * dt;
        src_seek(src, PH_STREAM_AUX, 0);
        double extdata1 = gain * latent + sigma1 * src_next(src);
        double choicert;
        if (ev >= bound) { choicert = ter + rt; o->choice = 1; }
        else if (ev <= 0) { choicert = -ter - rt; o->choice = -1; }
        else { choicert = 0; o->choice = 0; }
        o->out0 = choicert;
        o->out1 = extdata1;
        o->bound = bound;
    }
    o->evidence = ev;
    o->n_steps = (int64_t)n;
    o->consumed = (int64_t)(src->pos - pos0);
}

/* ------------------------------------------------------------------ */
/* Dataset-level entry points                                           */
/* ------------------------------------------------------------------ */
static int n_params_of(int model) {
    switch (model) {
    case ORC_MODEL_BASIC: return 5;
    case ORC_MODEL_ALPHA_SCALE: return 8;
    case ORC_MODEL_TRIALWISE: return 4;
    case ORC_MODEL_ETA: return 6;
    case ORC_MODEL_GENERAL: return 24;
    default: return 7;
    }
}

ORC_EXPORT int orc_n_params(int model) { return n_params_of(model); }

static int n_cols_of(int model) { return model == ORC_MODEL_GENERAL ? 3 : 2; }

ORC_EXPORT int orc_n_cols(int model) { return n_cols_of(model); }

static void store_trial(const orc_trial *t, size_t i, int ncols, double *out, int64_t *n_steps,
                        int32_t *choice, double *evidence, double *bound, int64_t *consumed) {
    out[ncols * i] = t->out0;
    out[ncols * i + 1] = t->out1;
    if (ncols > 2) out[ncols * i + 2] = t->out2;
    if (n_steps) n_steps[i] = t->n_steps;
    if (choice) choice[i] = t->choice;
    if (evidence) evidence[i] = t->evidence;
    if (bound) bound[i] = t->bound;
    if (consumed) consumed[i] = t->consumed;
}

/* simulate_trials(params, n_trials) with an explicit buffer of standard
 * normals consumed sequentially across trials, exactly as the reference's
 * single global stream is.  Returns 0, or -1 if the buffer ran out. */
ORC_EXPORT int orc_simulate_buffer(int model, const double *params, int64_t n_trials,
                                   const double *bound_in, double dt, double max_steps,
                                   int flags, const double *normals, int64_t n_normals,
                                   double *out, int64_t *n_steps, int32_t *choice,
                                   double *evidence, double *bound, int64_t *consumed) {
    orc_src src;
    memset(&src, 0, sizeof(src));
    src.kind = SRC_BUFFER;
    src.buf = normals;
    src.len = (size_t)n_normals;
    for (int64_t i = 0; i < n_trials; i++) {
        orc_trial t;
        double b = bound_in ? bound_in[i] : 0.0;
        if (model == ORC_MODEL_TRIALWISE && b < 0) return -2; /* ValueError in the reference */
        trial_run(model, params, b, dt, max_steps, flags, &src, &t);
        if (src.overrun) return -1;
        store_trial(&t, (size_t)i, n_cols_of(model), out, n_steps, choice, evidence, bound, consumed);
    }
    return 0;
}

/* Same, drawing from the MT19937/polar stream seeded like np.random.seed(seed)
 * inside numba-jitted code.  Reproduces the reference's own output. */
ORC_EXPORT int orc_simulate_mt(int model, const double *params, int64_t n_trials,
                               const double *bound_in, double dt, double max_steps,
                               int flags, uint32_t seed, double *out, int64_t *n_steps,
                               int32_t *choice, double *evidence, double *bound) {
    orc_src src;
    memset(&src, 0, sizeof(src));
    src.kind = SRC_MT;
    orc_mt_seed(&src.mt, seed);
    for (int64_t i = 0; i < n_trials; i++) {
        orc_trial t;
        double b = bound_in ? bound_in[i] : 0.0;
        if (model == ORC_MODEL_TRIALWISE && b < 0) return -2;
        trial_run(model, params, b, dt, max_steps, flags, &src, &t);
        store_trial(&t, (size_t)i, n_cols_of(model), out, n_steps, choice, evidence, bound, NULL);
    }
    return 0;
}

/* Same, drawing the fp64-ideal normals of the Philox counter stream the
 * CUDA kernels use: trial i of dataset `dataset` reads counters
 * (block, trial_offset+i, dataset, stream). */
ORC_EXPORT int orc_simulate_philox(int model, const double *params, int64_t n_trials,
                                   const double *bound_in, double dt, double max_steps,
                                   int flags, uint64_t seed, uint32_t dataset,
                                   uint32_t trial_offset, double *out, int64_t *n_steps,
                                   int32_t *choice, double *evidence, double *bound) {
    orc_src src;
    memset(&src, 0, sizeof(src));
    src.kind = SRC_PHILOX;
    src.seed = seed;
    src.dataset = dataset;
    for (int64_t i = 0; i < n_trials; i++) {
        orc_trial t;
        double b = bound_in ? bound_in[i] : 0.0;
        if (model == ORC_MODEL_TRIALWISE && b < 0) return -2;
        src.trial = trial_offset + (uint32_t)i;
        trial_run(model, params, b, dt, max_steps, flags, &src, &t);
        store_trial(&t, (size_t)i, n_cols_of(model), out, n_steps, choice, evidence, bound, NULL);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Evidence-path variants (retired_models/)      */
/*   basic_ddm_dc_evidence.py:87-151        200 obs, noise sigma1, per-trial z-score  (mode 1)  */
/*   basic_ddm_dc_evidence2.py:83-150       200 obs, noise sigma1, dataset-level
 *                                          (x - mean(path_means)) / std(path_means) (mode 2)  */
/*   basic_ddm_dc_evidence_no_noise2.py:82-147  400 obs, noise .001, per-trial z-score (mode 1) */
/* p = [drift, boundary, beta, tau, dc, sigma1]; row = (rt, choice, path[n_obs]).              */
/* Normals per trial, in order: z_step[0..n-1], z_noise[0..n_obs-1].  Means and variances are  */
/* plain left-to-right sums (numba's array_mean / array_var), std = sqrt(var).                 */
/* ------------------------------------------------------------------ */
static double seq_mean(const double *v, int64_t n) {
    double c = 0.0;
    for (int64_t i = 0; i < n; i++) c += v[i];
    return c / (double)n;
}

static double seq_std(const double *v, int64_t n) {
    double m = seq_mean(v, n), ssd = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double d = v[i] - m;
        ssd += d * d;
    }
    return sqrt(ssd / (double)n);
}

static void evidence_trial(const double *p, int n_obs, int mode, double dt, double max_steps, orc_src *src,
                           double *row, int64_t *n_out, double *mean_out) {
    double drift = p[0], boundary = p[1], beta = p[2], tau = p[3], dc = p[4], sigma1 = p[5];
    double n_steps = 0.0, evidence = boundary * beta;
    double *path = row + 2;
    for (int k = 0; k < n_obs; k++) path[k] = 0.0;
    src_seek(src, PH_STREAM_STEP, 0);
    while ((evidence > 0) && (evidence < boundary) && (n_steps < max_steps)) {
        double z = src_next(src);
        double t1 = drift * dt;
        double t2 = sqrt(dt) * dc;
        double t3 = t2 * z;
        evidence = evidence + (t1 + t3);
        if (n_steps < (double)n_obs) path[(int)n_steps] = evidence;
        n_steps += 1.0;
    }
    row[0] = n_steps * dt + tau;
    for (int k = (int)n_steps; k < n_obs; k++) path[k] = evidence; /* held at the final value */
    src_seek(src, PH_STREAM_AUX, 0);
    for (int k = 0; k < n_obs; k++) path[k] = path[k] + (0.0 + sigma1 * src_next(src));
    if (mode == 1) {
        double m = seq_mean(path, n_obs), sd = seq_std(path, n_obs);
        for (int k = 0; k < n_obs; k++) path[k] = (path[k] - m) / sd;
    } else if (mean_out) {
        *mean_out = seq_mean(path, n_obs);
    }
    row[1] = (evidence >= boundary) ? 1.0 : ((evidence <= 0) ? -1.0 : 0.0);
    *n_out = (int64_t)n_steps;
}

/* src_kind: 0 buffer (normals, n_normals), 1 MT19937 (seed), 2 Philox (seed, dataset, trial_offset). */
ORC_EXPORT int orc_simulate_evidence(const double *params, int64_t n_trials, int n_obs, int mode, double dt,
                                     double max_steps, int src_kind, const double *normals, int64_t n_normals,
                                     uint64_t seed, uint32_t dataset, uint32_t trial_offset, double *out,
                                     int64_t *n_steps, int64_t *consumed) {
    if (n_obs < 1 || mode < 0 || mode > 2) return -3;
    orc_src src;
    memset(&src, 0, sizeof(src));
    src.kind = src_kind;
    if (src_kind == SRC_BUFFER) { src.buf = normals; src.len = (size_t)n_normals; }
    else if (src_kind == SRC_MT) orc_mt_seed(&src.mt, (uint32_t)seed);
    else { src.seed = seed; src.dataset = dataset; }
    const int64_t cols = 2 + n_obs;
    double *means = (mode == 2) ? (double *)malloc(sizeof(double) * (size_t)(n_trials > 0 ? n_trials : 1)) : NULL;
    for (int64_t i = 0; i < n_trials; i++) {
        size_t pos0 = src.pos;
        int64_t n;
        src.trial = trial_offset + (uint32_t)i;
        evidence_trial(params, n_obs, mode, dt, max_steps, &src, out + (size_t)i * cols, &n, means ? &means[i] : NULL);
        if (src.overrun) { free(means); return -1; }
        if (n_steps) n_steps[i] = n;
        if (consumed) consumed[i] = (int64_t)(src.pos - pos0);
    }
    if (mode == 2 && n_trials > 0) {
        double m = seq_mean(means, n_trials), sd = seq_std(means, n_trials);
        for (int64_t i = 0; i < n_trials; i++)
            for (int k = 0; k < n_obs; k++) {
                double *v = out + (size_t)i * cols + 2 + k;
                *v = (*v - m) / sd;
            }
    }
    free(means);
    return 0;
}

/* ------------------------------------------------------------------ */
/* CPU baseline: B datasets x n_trials on `n_threads` host threads, each */
/* dataset on its own MT19937 stream (seed + dataset), i.e. the          */
/* reference's numba loop run process-parallel over datasets.  Returns   */
/* total Euler steps executed via *total_steps.                          */
/* ------------------------------------------------------------------ */
typedef struct {
    int model, flags;
    const double *params;
    int n_params;
    int64_t n_datasets, n_trials;
    double dt, max_steps;
    uint32_t seed;
    double *out;
    int64_t next;           /* shared dataset cursor */
    pthread_mutex_t lock;
    int64_t total_steps, total_timeouts;
} bench_job;

static void *bench_worker(void *arg) {
    bench_job *job = (bench_job *)arg;
    int64_t steps = 0, timeouts = 0;
    double *scratch = NULL;
    if (!job->out) scratch = (double *)malloc(sizeof(double) * 3 * (size_t)job->n_trials);
    for (;;) {
        pthread_mutex_lock(&job->lock);
        int64_t d = job->next++;
        pthread_mutex_unlock(&job->lock);
        if (d >= job->n_datasets) break;
        orc_src src;
        memset(&src, 0, sizeof(src));
        src.kind = SRC_MT;
        orc_mt_seed(&src.mt, job->seed + (uint32_t)d);
        const double *p = job->params + (size_t)d * job->n_params;
        const int nc = n_cols_of(job->model);
        double *o = job->out ? job->out + (size_t)d * nc * job->n_trials : scratch;
        for (int64_t i = 0; i < job->n_trials; i++) {
            orc_trial t;
            trial_run(job->model, p, 0.0, job->dt, job->max_steps, job->flags, &src, &t);
            o[nc * i] = t.out0;
            o[nc * i + 1] = t.out1;
            if (nc > 2) o[nc * i + 2] = t.out2;
            steps += t.n_steps;
            timeouts += (t.choice == 0);
        }
    }
    free(scratch);
    pthread_mutex_lock(&job->lock);
    job->total_steps += steps;
    job->total_timeouts += timeouts;
    pthread_mutex_unlock(&job->lock);
    return NULL;
}

ORC_EXPORT int orc_simulate_batch_mt(int model, const double *params, int64_t n_datasets,
                                     int64_t n_trials, double dt, double max_steps, int flags,
                                     uint32_t seed, int n_threads, double *out,
                                     int64_t *total_steps, int64_t *total_timeouts) {
    if (model == ORC_MODEL_TRIALWISE) return -3;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    bench_job job;
    memset(&job, 0, sizeof(job));
    job.model = model; job.flags = flags; job.params = params;
    job.n_params = n_params_of(model);
    job.n_datasets = n_datasets; job.n_trials = n_trials;
    job.dt = dt; job.max_steps = max_steps; job.seed = seed; job.out = out;
    pthread_mutex_init(&job.lock, NULL);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    int started = 0;
    for (int i = 0; i < n_threads; i++)
        if (pthread_create(&th[i], NULL, bench_worker, &job) == 0) started++; else break;
    if (started == 0) bench_worker(&job);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&job.lock);
    if (total_steps) *total_steps = job.total_steps;
    if (total_timeouts) *total_timeouts = job.total_timeouts;
    return 0;
}
