/*
 * ddm_b200.h -- C ABI of the B200-native drift-diffusion trial simulator.
 *
 * Drop-in boundary for the data-parallel hot path of mdnunez/bayesflow_nddms.
 * The reference has no FFI of its own (it is pure Python + numba); the operator
 * interface this library sits behind is BayesFlow 1.1's simulation API as the
 * reference uses it (basic_ddm_dc.py:130-134).  Each entry point below names
 * the reference callable it replaces.  Python binds these with ctypes
 * (bayesflow_nddms_b200/_capi.py); INTEGRATION.md shows the stub.
 *
 * Conventions: every function returns 0 (DDM_OK) or a negative ddm_status and
 * never throws or exits; ddm_last_error(ctx) holds the message.  One ctx per
 * host thread / rank; functions on distinct ctx are concurrency-safe; there is
 * no global mutable state.  Host pointers are plain C arrays; no torch types.
 * Device work is enqueued on the ctx stream (ddm_set_stream to borrow one).
 */
#ifndef DDM_B200_H
#define DDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDM_B200_VERSION 200

typedef struct ddm_ctx ddm_ctx;
struct DLManagedTensor; /* include/ddm_dlpack.h (DLPack v0.8 ABI) */

/* Simulator variants.  Parameter vectors keep the reference's order. */
enum ddm_model {
    /* basic_ddm_dc.py:85-125.  params[5] = drift, boundary, beta, tau, dc.
       out = (rt = n*dt + tau, choice in {+1,-1, 0|1 on timeout}) */
    DDM_MODEL_BASIC = 0,
    /* single_trial_alpha_not_scaled.py:107-155.  params[7] = drift, mu_alpha,
       beta, ter, std_alpha, dc, sigma1.  out = (signed choicert, extdata1) */
    DDM_MODEL_ALPHA = 1,
    /* :926-974 (diffusion_trial_alt): params[7] = drift, alpha, beta, ter,
       std_dc, mu_dc, sigma1; per-trial dc; extdata1 = N(dc_trial, sigma1) */
    DDM_MODEL_ALPHA_DC = 2,
    /* :1237-1285 (_scale): params[8] = M1 + gamma; extdata1 = N(gamma*bound, sigma1) */
    DDM_MODEL_ALPHA_SCALE = 3,
    /* :1471-1519 (_scale2): params[7]; extdata1 = N(2*bound, sigma1) */
    DDM_MODEL_ALPHA_SCALE2 = 4,
    /* imputation_from_stahl_not_scaled.py:120-148: per-trial supplied boundary,
       per-group params[4] = drift, beta, ter, dc.  out = (signed choicert, bound) */
    DDM_MODEL_TRIALWISE = 5,
    /* retired_models/basic_ddm_eta_dc.py:80-120: per-trial drift ~ N(mu_drift, eta) (one pre-draw, no
       rejection).  params[6] = mu_drift, alpha, beta, ter, eta, dc.  out = (rt, choice) as DDM_MODEL_BASIC */
    DDM_MODEL_ETA = 6,
    /* The retired zoo's two-latent / two-channel simulators in one parametrisation, e.g.
       retired_models/single_trial_drift_dc5.py:90-154, single_trial_drift_dc4.py:90-146,
       single_trial_alpha_dc.py:109-176.  params[24] =
         0 drift_mu 1 drift_sd | 2 bound_mu 3 bound_sd | 4 dc_mu 5 dc_sd | 6 beta 7 tau
         8..13  channel 1: coefficient of drift_t, bound_t, dc_t; sigma; shift; scale
                ext = ((c_d*drift_t + c_b*bound_t) + c_dc*dc_t + sigma*z - shift) / scale
         14..19 channel 2 likewise | 20 order of the pre-draws (index into the permutations of
         (drift, bound, dc): 0 = d,b,c  1 = d,c,b  2 = b,d,c  3 = b,c,d  4 = c,d,b  5 = c,b,d; it matters
         only in shared-increment mode) | 21 number of channels | 22 output style | 23 reserved.
       A latent with sd == 0 is fixed; drift is drawn once; boundary and dc are redrawn until > 0.
       THREE output columns: style 0 (rt, choice, ext1), style 1 (signed choicert, ext1, ext2).
       bayesflow_nddms_b200/two_channel.py maps the reference's parameter vectors onto these. */
    DDM_MODEL_GENERAL = 7
};

/* Prior families of ddm_draw_prior; the value is the ddm_model whose parameter layout is produced
 * (7 and 8 have no model id of their own). */
enum ddm_prior {
    DDM_PRIOR_BASIC = 0,        /* basic_ddm_dc.py:62-80 -> (drift, alpha, beta, ter, dc) */
    DDM_PRIOR_ALPHA = 1,        /* single_trial_alpha_not_scaled.py:78-102 -> 7 columns */
    DDM_PRIOR_ALPHA_DC = 2,     /* :899-923 (draw_prior_alt) -> 7 columns */
    DDM_PRIOR_ALPHA_SCALE = 3,  /* :1205-1232 (draw_prior_scale) -> 8 columns */
    DDM_PRIOR_ALPHA_SCALE2 = 4, /* the model's own prior, 7 columns */
    DDM_PRIOR_ETA = 6,          /* retired_models/basic_ddm_eta_dc.py:54-75 -> 6 columns */
    DDM_PRIOR_SWEEP = 7,        /* the basic prior with ter = 0 (throughput sweep) */
    DDM_PRIOR_EVIDENCE = 8      /* retired_models/basic_ddm_dc_evidence.py:61-82 -> 6 columns (…, dc, sigma1) */
};

enum ddm_status {
    DDM_OK = 0,
    DDM_ERR_INVALID = -1,        /* bad argument */
    DDM_ERR_CUDA = -2,           /* CUDA runtime error (message in ddm_last_error) */
    DDM_ERR_NOMEM = -3,
    DDM_ERR_NEGATIVE_BOUND = -4, /* the reference raises ValueError (imputation...:124-125) */
    DDM_ERR_STATE = -5           /* call order (e.g. download before run) */
};

enum ddm_flags {
    /* basic_ddm_dc.py:110-112 leaves `choice` unbound on a timeout; under numba
       0.65 the trial then reports choice = 1.  Default here is 0 (the author's
       stated intent, and what every other variant does); this flag reproduces
       the numba artefact. */
    DDM_FLAG_TIMEOUT_CHOICE_ONE = 1,
    DDM_FLAG_OUT_F32 = 2,      /* outputs are float32 pairs (configurator dtype) instead of float64 */
    DDM_FLAG_KEEP_STEPS = 4,   /* also record int32 Euler-step counts per trial */
    DDM_FLAG_FORCE_GENERIC = 8, /* use the one-thread-per-trial kernel (validation) */
    DDM_FLAG_OUT_STATE = 16,    /* validation: column 1 := the trial's final evidence (reference frame) */
    DDM_FLAG_F32_NORMALS = 32   /* validation, precision 64: run the reference's fp64 arithmetic on the fp32 production
                                   kernels' normals (the "exported increments" of check #1, at scale, on the device) */
};

typedef struct ddm_stats {
    uint64_t n_trials;       /* trials simulated by the last run */
    uint64_t total_steps;    /* Euler steps executed (sum of per-trial n) */
    uint64_t n_timeouts;     /* trials that hit max_steps */
    uint64_t n_upper;        /* trials absorbed at the upper boundary */
    uint64_t reject_cap_hits;/* per-trial redraw loops that hit the iteration cap */
    double kernel_ms;        /* CUDA-event time of the simulator kernel(s), launching stream */
    int32_t kernel_launches; /* kernels of this library launched by the last run */
    int32_t used_persistent; /* 1 if a production fp32 kernel ran (scheduler 1-3), 0 if the generic / validation kernel */
    int32_t grid, block, refill_threshold, tile;
    uint64_t debug_overruns; /* shared-increment mode: trials that ran past the normals buffer */
    uint64_t d2h_bytes;      /* output bytes the run copied device -> host itself (0: batch left resident) */
    int32_t host_decode_threads; /* > 0: compact wire records were expanded by that many host threads */
    int32_t scheduler;       /* 0 one thread per trial (generic / validation), 1 round-1 persistent kernel, 2 tile kernel,
                              * 3 latency kernel (small launches) */
} ddm_stats;

/* ---- lifecycle -------------------------------------------------------- */
int ddm_version(void);
/* Philox4x32 rounds this library was built with: 10 (shipped), or 7 for the measurement-only variant
 * libddm_b200_philox7.so (bayesflow_nddms_b200/_build.py: build_philox7). */
int ddm_philox_rounds(void);
int ddm_create(int device, ddm_ctx **out);
int ddm_destroy(ddm_ctx *ctx);
const char *ddm_last_error(const ddm_ctx *ctx); /* ctx may be NULL: last create error */
int ddm_set_stream(ddm_ctx *ctx, void *cuda_stream); /* NULL restores the ctx-owned stream */
int ddm_synchronize(ddm_ctx *ctx);
/* tuning knobs; 0 = automatic */
int ddm_set_tuning(ddm_ctx *ctx, int refill_threshold, int blocks_per_sm, int tile);
/* Scheduler of the production (precision 32) path: 0 the tile-staged persistent kernel (set-up and emission a tile at a
 * time through shared memory, finished lanes refilled inside the stepping loop); 1 the round-1 persistent kernel
 * (per-lane set-up and emission inside a refill pass), kept for A/B measurements; 2 the latency kernel (one thread per
 * trial, speculative six-step blocks with the next blocks' normals drawn under them: a launch with nothing to refill
 * lasts as long as its longest trial's dependent chain); -1, the default: the latency kernel for launches of at most
 * 256 Ki trials -- the reference's own batch sizes -- and the tile kernel above that.  Results are bit-identical. */
int ddm_set_kernel_variant(ddm_ctx *ctx, int variant);
/* ddm_simulate with a host destination streams batches of at least min_rows trials to the host in
 * chunks of about chunk_rows trials, overlapping kernel and PCIe copy (defaults: min_rows 8 Mi trials, 10^6 into a
 * page-locked destination; 10^6 -- 4 Mi for the layouts with an external column -- when the rows travel as compact
 * records, see ddm_set_host_decode; chunk_rows: the smaller of a
 * quarter of the batch and half of what is left, within 2 Mi .. 32 Mi -- 512 Ki for batches below 4 Mi --
 * so that the last chunks are small; values < 0 restore them -- chunk_rows -2 / -3 select "a quarter" / "half of what is left"
 * alone, for measurements; a huge min_rows switches the pipeline off).  After such a run the batch
 * is not resident on the device.  Results do not depend on the chunking. */
int ddm_set_pipeline(ddm_ctx *ctx, int64_t min_rows, int64_t chunk_rows);
/* Two-column models cross PCIe in that pipeline as compact records -- (Euler steps, choice) in 4 bytes,
 * plus the fp32 external measurement for the single-trial-boundary models, instead of 16-byte float64
 * rows -- and n_threads host threads (0 = automatic: the process's CPUs / GPUs on the box, at most 16)
 * write the rows of basic_ddm_dc.py:108-112 (rt = n*dt + ndt, choice) into out_host while the next chunk
 * is simulated.  Same bits as the plain copy; n_threads < 0 switches back to float64 rows over PCIe. */
int ddm_set_host_decode(ddm_ctx *ctx, int n_threads);

/* ---- the hot path ------------------------------------------------------ */
/* Replaces B calls of simulate_trials(params[b], n_trials)  (basic_ddm_dc.py:114-125,
 * single_trial_alpha_not_scaled.py:144-155 and the _alt/_scale/_scale2/_fine variants),
 * i.e. BayesFlow's batch_simulator_fun contract: params (B,P) f64 host ->
 * out_host (B, n_trials, 2) f64 (or f32 with DDM_FLAG_OUT_F32); (B, n_trials, 3) for DDM_MODEL_GENERAL.  out_host may be
 * NULL: results stay on the device (ddm_last_output_dlpack / ddm_download).
 * dt, max_steps: the reference's default kwargs (.01, 400).  Philox key = seed;
 * counter words = (stream, step block, trial, dataset_offset + b), so the result of
 * dataset b does not depend on how datasets are sharded over GPUs.  dataset_offset is a
 * 64-bit global index below 2^56: its low 32 bits are a counter word, the rest rides in the
 * stream word, so a long run rolls over into fresh counters; one call must not straddle a
 * multiple of 2^32 (DDM_ERR_INVALID: start the batch at the next multiple).
 * precision: 32 (production) or 64 (validation; reference operation order). */
int ddm_simulate(ddm_ctx *ctx, int model, const double *params, int64_t n_datasets, int n_params,
                 int64_t n_trials, double dt, int max_steps, uint64_t seed, uint64_t dataset_offset,
                 int precision, int flags, void *out_host);

/* Replaces B calls of draw_prior() (see enum ddm_prior): one launch writes the (n_draws, P) float64
 * parameter matrix into the context's parameter arena -- a following ddm_run simulates these draws
 * with no host round trip -- and, if params_host != NULL, copies it out (BayesFlow's 'prior_draws').
 * Draw i uses Philox counters keyed by draw_offset + i, so draws are reproducible and shardable. */
int ddm_draw_prior(ddm_ctx *ctx, int prior, int64_t n_draws, uint64_t seed, uint64_t draw_offset, double *params_host);

/* The same in three steps, for callers that keep inputs/outputs resident. */
int ddm_upload_params(ddm_ctx *ctx, int model, const double *params, int64_t n_datasets, int n_params);
int ddm_run(ddm_ctx *ctx, int64_t n_trials, double dt, int max_steps, uint64_t seed,
            uint64_t dataset_offset, int precision, int flags);
int ddm_download(ddm_ctx *ctx, void *out_host); /* dtype as chosen by the run's flags */

/* Replaces the per-row loop imputation_from_stahl_not_scaled.py:205-213:
 * trial i uses bound[i] and group_params[group[i]] = (drift, beta, ter, dc).
 * out_host (n, 2): col0 signed choicert, col1 the boundary used (>= 0).
 * Any bound[i] < 0 -> DDM_ERR_NEGATIVE_BOUND, nothing is simulated.  Philox counters
 * are (step block, trial_offset + i, 0, stream): shard by trial range with trial_offset. */
int ddm_simulate_trialwise(ddm_ctx *ctx, const int32_t *group, const double *bound,
                           const double *group_params, int64_t n, int n_groups, double dt,
                           int max_steps, uint64_t seed, uint64_t trial_offset, int precision,
                           int flags, void *out_host);

/* Evidence-path variants of the retired model zoo (retired_models/basic_ddm_dc_evidence.py:87-151,
 * basic_ddm_dc_evidence2.py:83-150, basic_ddm_dc_evidence_no_noise2.py:82-147): B calls of
 * simulate_trials(params[b], n_trials) -> out_host (B, n_trials, 2 + n_obs): rt, choice, then the
 * first n_obs evidence values of the path (held at the final evidence after the crossing) plus
 * N(0, sigma1) noise, standardised.  params (B,6) = drift, boundary, beta, tau, dc, sigma1 (the
 * no-noise variants pass sigma1 = .001).  n_obs = int(.2/dt) or int(.4/dt) in the reference.
 * standardize: 0 none, 1 per-trial z-score (evidence, no_noise*), 2 dataset-level
 * (x - mean(path_means)) / std(path_means) (evidence2).  Noise normals come from the aux Philox
 * stream of the trial (index k).  Shared-increment mode needs precision 64. */
int ddm_simulate_evidence(ddm_ctx *ctx, const double *params, int64_t n_datasets, int64_t n_trials, int n_obs,
                          int standardize, double dt, int max_steps, uint64_t seed, uint64_t dataset_offset,
                          int precision, int flags, void *out_host);

/* Replaces B calls of simulratcliff(N, Alpha, Tau, Nu, Beta, rangeTau, rangeBeta, Eta, Varsigma)
 * (pyhddmjagsutils.py:47-176; caller alpha_not_scaled.py:95-97): the exact rejection sampler of Tuerlinckx et al.
 * (2001) -- no time step -- with per-trial drift N(Nu, Eta), start point and non-decision time ranges.
 * params (B, 8) f64 host, columns [Alpha, Tau, Nu, Beta, rangeTau, rangeBeta, Eta, Varsigma];
 * out_host (B, n_trials) f64 signed response times (+ upper boundary, - lower), may be NULL (resident).
 * |Nu| is clipped to 5 and Eta == 0 replaced by 1e-16 as in the reference.  Stats: total_steps counts
 * the symmetric intervals walked, n_upper the upper-boundary responses. */
int ddm_simulate_exact(ddm_ctx *ctx, const double *params, int64_t n_datasets, int64_t n_trials, uint64_t seed,
                       uint64_t dataset_offset, double *out_host);

/* Per-trial Euler-step counts of the last run (needs DDM_FLAG_KEEP_STEPS). */
int ddm_last_steps(ddm_ctx *ctx, int32_t *steps_host);
int ddm_last_stats(ddm_ctx *ctx, ddm_stats *out);
/* One online-training batch in one call -- replaces bf.simulation.GenerativeModel(prior, simulator)(batch_size)
 * (basic_ddm_dc.py:129-133; single_trial_alpha_not_scaled.py:270-274): ddm_draw_prior on the device (draws keyed by
 * draw_offset ..), ddm_run on the resident draws (datasets keyed by the same indices; precision 32, `flags` as in
 * ddm_run), the (n_draws, P) draws copied to params_host (may be NULL), the batch handed over like
 * ddm_last_output_dlpack.  Same bits as the three calls; one stream synchronisation instead of two. */
int ddm_training_batch(ddm_ctx *ctx, int prior, int64_t n_draws, int64_t n_trials, double dt, int max_steps, uint64_t seed,
                       uint64_t draw_offset, int flags, double *params_host, struct DLManagedTensor **out);

/* Response-time histogram of the last run's resident output, reduced on the device (no row leaves HBM):
 * hist_host[0 .. n_bins) counts upper-boundary responses with |rt| in [k, k+1) * rt_max / n_bins,
 * hist_host[n_bins .. 2 n_bins) the lower-boundary ones, hist_host[2 n_bins] trials without a response
 * (timeouts: choice 0 / signed rt 0) and hist_host[2 n_bins + 1] responses with |rt| >= rt_max.
 * n_bins <= 8192.  Additive over datasets: shards and GPUs sum.  Not for DDM_MODEL_GENERAL (mixed layouts). */
int ddm_last_output_histogram(ddm_ctx *ctx, int n_bins, double rt_max, uint64_t *hist_host);

/* The throughput sweep as SURVEY.md section 8d specifies it ("outputs reduced on device -- no host copy"): one
 * call takes host parameters, simulates the batch into resident float32 rows, reduces them on the device and
 * returns only the histogram of ddm_last_output_histogram (2 n_bins + 2 counters).  Replaces
 * [simulate_trials(p, n_trials) for p in params] followed by a histogram of the stacked rows; two-column
 * dataset-wise models only.  The rows stay resident afterwards (ddm_last_output_dlpack / ddm_download).  From 64 Mi
 * trials on the batch is produced in a few chunks of datasets (ddm_histogram_chunks) -- each chunk's parameters cross
 * PCIe while the previous chunk is simulated and its rows are reduced beside the next chunk's kernel -- into the same
 * resident buffer and with the same results (ddm_set_pipeline's min_rows / chunk_rows govern the threshold and the
 * smallest chunk here, too). */
int ddm_simulate_histogram(ddm_ctx *ctx, int model, const double *params, int64_t n_datasets, int n_params,
                           int64_t n_trials, double dt, int max_steps, uint64_t seed, uint64_t dataset_offset,
                           int precision, int flags, int n_bins, double rt_max, uint64_t *hist_host);

/* Device hand-off of the last run's output as a DLPack tensor of shape
 * (n_datasets, n_trials, 2) ((n, 2) for trialwise, (n_datasets, n_trials, 2 + n_obs) for evidence runs), dtype per the run's flags,
 * on this ctx's device.  Ownership of the buffer moves to the consumer; its
 * deleter frees it.  The producer stream is synchronised before return. */
int ddm_last_output_dlpack(ddm_ctx *ctx, struct DLManagedTensor **out);
/* Raw device pointer of the last output (borrowed; valid until the next run). */
int ddm_last_output_device_ptr(ddm_ctx *ctx, void **ptr, size_t *bytes);

/* ---- parity / validation hooks ----------------------------------------- */
/* Shared-increment mode: subsequent runs read standard normals from this buffer
 * instead of Philox.  Trial t (flat index dataset*n_trials + trial) consumes
 * z[offsets[t]], z[offsets[t]+1], ... in the reference's order (pre-draws, steps,
 * ext).  z == NULL switches the mode off.  Forces the generic kernel. */
int ddm_set_normals_debug(ddm_ctx *ctx, const double *z, size_t n, const int64_t *offsets, int64_t n_trials);
/* The normals the production (precision 32) or validation (64) kernels use for
 * counters (dataset, trial, stream), indices first..first+count-1, as doubles. */
int ddm_export_normals(ddm_ctx *ctx, uint64_t seed, uint64_t dataset, uint32_t trial, uint32_t stream,
                       uint32_t first, uint32_t count, int precision, double *out_host);
/* The production normal generator under the microscope: ceil(n_normals / 6) Philox blocks through the fp32 map of the
 * stepping kernels (21-bit Box-Muller fields, MUFU lg2/sqrt/sin/cos), reduced on the device.  hist_host receives
 * n_bins_abs counts of |z| in equal bins over [0, z_max), one count of |z| >= z_max, then n_bins_angle counts of the
 * Box-Muller pairs' angles in equal sectors; moments_host[4] = sum z, z^2, z^3, z^4. */
int ddm_normals_histogram(ddm_ctx *ctx, uint64_t seed, uint64_t n_normals, int n_bins_abs, double z_max, int n_bins_angle,
                          uint64_t *hist_host, double *moments_host);
/* Raw Philox4x32-10 blocks computed on the device (known-answer tests). */
int ddm_philox4x32(ddm_ctx *ctx, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4, int64_t n_blocks);
/* The chunk schedule ddm_simulate's streamed path uses for a batch (no GPU needed): writes up to `capacity`
 * (first dataset, datasets) pairs and returns the number of chunks.  chunk_rows as in ddm_set_pipeline. */
int64_t ddm_pipeline_chunks(int64_t n_datasets, int64_t n_trials, int64_t chunk_rows, int64_t *first, int64_t *count,
                            int64_t capacity);
/* The chunk schedule ddm_simulate_histogram uses from 64 Mi trials on (no GPU needed): a sixteenth of the batch, then
 * half of what is left each time, no chunk below min_chunk_rows trials (<= 0: the default, 32 Mi). */
int64_t ddm_histogram_chunks(int64_t n_datasets, int64_t n_trials, int64_t min_chunk_rows, int64_t *first, int64_t *count,
                             int64_t capacity);
/* The host half of ddm_set_host_decode on its own (no GPU needed): expands n_datasets * n_trials wire
 * records -- int32 (steps << 2 | choice + 1) when basic_columns, else {that, fp32 bits} pairs -- into
 * (rows, 2) float64 / float32 with n_threads threads.  tau = params[d * n_params + 3]. */
int ddm_wire_decode_host(const void *wire, void *out_host, const double *params, int n_params, int64_t n_datasets,
                         int64_t n_trials, double dt, int basic_columns, int flags, int n_threads);

/* ---- measurement -------------------------------------------------------- */
/* Pipe micro-benchmarks for the issue roofline.  which: see ddm_microbench_id.
 * Returns warp-instructions (of the measured kind) per second chip-wide in
 * *inst_per_s and the SM clock seen (cycles/s from clock64 over event time). */
enum ddm_microbench_id {
    DDM_MB_FFMA = 0, DDM_MB_IMAD_WIDE = 1, DDM_MB_LOP3 = 2, DDM_MB_IADD3 = 3,
    DDM_MB_MUFU_LG2 = 4, DDM_MB_MUFU_SIN = 5, DDM_MB_MIX_FMA_ALU = 6, DDM_MB_FSETP = 7,
    DDM_MB_PHILOX = 8, DDM_MB_NORMALS = 9, DDM_MB_MIX_IMADW_LOP3 = 10, DDM_MB_MIX_MUFU_LOP3 = 11,
    DDM_MB_MIX_MUFU_IMADW = 12, DDM_MB_MIX_BLOCKLIKE = 13,
    /* round 2: register-operand forms, the Philox round's multiply, packed fp32, 7 rounds, two trials per lane */
    DDM_MB_FFMA_REG = 14, DDM_MB_FADD_REG = 15, DDM_MB_IMAD_WIDE_NOACC = 16, DDM_MB_IMAD_HI = 17, DDM_MB_FFMA2 = 18,
    DDM_MB_PHILOX7 = 19, DDM_MB_NORMALS2 = 20, DDM_MB_COUNT = 21
};
int ddm_microbench(ddm_ctx *ctx, int which, int iters, double *inst_per_s, double *sm_hz);

/* The ceiling of the host side of ddm_set_host_decode on this box: n_threads host threads (0 = the automatic
 * count) fill `bytes` of host memory with the decode's own non-temporal stores; *bytes_per_s = the best of five
 * passes.  No GPU needed. */
int ddm_host_stream_peak(int n_threads, size_t bytes, double *bytes_per_s);

/* pinned host memory helpers (for callers that want full-rate D2H) */
int ddm_host_alloc(size_t bytes, void **ptr);
int ddm_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* DDM_B200_H */
