// ddm_capi.cu -- the C ABI of include/ddm_b200.h over the kernels in ddm_kernels.cu.
//
// One ddm_ctx owns a device, a stream, grow-only device arenas and a pool of output
// buffers (so that a steady-state online-training loop performs no cudaMalloc, even
// when every batch is handed away through DLPack).  No torch types, no global state
// except the last-create error string.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ddm_b200.h"
#include "../../include/ddm_dlpack.h"
#include "ddm_kernels.cuh"
#include "ddm_microbench.cuh"

#define DDM_API extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_create_error;

// ---- pooled device buffers (outputs can outlive the ctx via DLPack) -----------------
struct BufferPool {
    std::mutex mu;
    std::vector<std::pair<void *, size_t>> free_list;
    std::atomic<int> refs{1};
    int device = 0;
    size_t cached_bytes = 0;
    static constexpr size_t kMaxCachedBuffers = 4;

    void *take(size_t bytes, size_t *cap) {
        std::lock_guard<std::mutex> g(mu);
        int best = -1;
        for (int i = 0; i < (int)free_list.size(); i++)
            if (free_list[i].second >= bytes && (best < 0 || free_list[i].second < free_list[best].second)) best = i;
        if (best < 0 || free_list[best].second > 2 * bytes + (1u << 20)) return nullptr;
        void *p = free_list[best].first;
        *cap = free_list[best].second;
        free_list.erase(free_list.begin() + best);
        return p;
    }
    void give(void *p, size_t cap) {
        void *drop = nullptr;
        {
            std::lock_guard<std::mutex> g(mu);
            free_list.emplace_back(p, cap);
            if (free_list.size() > kMaxCachedBuffers) {
                // over capacity: let the largest go (a multi-GB buffer of an earlier, bigger batch is what costs memory,
                // and what a small-batch loop will never take again)
                size_t big = 0;
                for (size_t i = 1; i < free_list.size(); i++)
                    if (free_list[i].second > free_list[big].second) big = i;
                drop = free_list[big].first;
                free_list.erase(free_list.begin() + big);
            }
        }
        if (drop) {
            int prev = 0;
            cudaGetDevice(&prev);
            cudaSetDevice(device);
            cudaFree(drop);
            cudaSetDevice(prev);
        }
    }
    void release() {
        if (refs.fetch_sub(1) == 1) {
            int prev = 0;
            cudaGetDevice(&prev);
            cudaSetDevice(device);
            for (auto &b : free_list) cudaFree(b.first);
            cudaSetDevice(prev);
            delete this;
        }
    }
};

struct DlpackHolder {
    DLManagedTensor t;
    int64_t shape[3];
    void *buf;
    size_t cap;
    BufferPool *pool;
};

void dlpack_deleter(DLManagedTensor *self) {
    if (!self) return;
    auto *h = static_cast<DlpackHolder *>(self->manager_ctx);
    h->pool->give(h->buf, h->cap);
    h->pool->release();
    delete h;
}

template <typename T>
struct Arena {  // grow-only device array
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void free_() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

int n_params_of(int model) {
    switch (model) {
    case DDM_MODEL_BASIC: return 5;
    case DDM_MODEL_ALPHA_SCALE: return 8;
    case DDM_MODEL_TRIALWISE: return 4;
    case DDM_MODEL_ETA: return 6;
    case DDM_MODEL_GENERAL: return 24;
    case DDM_MODEL_ALPHA:
    case DDM_MODEL_ALPHA_DC:
    case DDM_MODEL_ALPHA_SCALE2: return 7;
    default: return -1;
    }
}

int n_cols_of(int model) { return model == DDM_MODEL_GENERAL ? 3 : 2; }

int kind_of(int model) {
    switch (model) {
    case DDM_MODEL_BASIC: return ddm::KIND_FIXED;
    case DDM_MODEL_ALPHA_DC: return ddm::KIND_DC;
    case DDM_MODEL_TRIALWISE: return ddm::KIND_TRIALWISE;
    case DDM_MODEL_ETA: return ddm::KIND_DRIFT;
    case DDM_MODEL_GENERAL: return ddm::KIND_GENERAL;
    default: return ddm::KIND_BOUND;
    }
}

}  // namespace

constexpr int kRunModelExact = 101;     // ... of ddm_simulate_exact runs ((B, N) signed response times)
constexpr int kRunModelEvidence = 100;  // run_model of ddm_simulate_evidence runs ((rt, choice, path) rows)

struct ddm_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    BufferPool *pool = nullptr;

    Arena<double> params;
    Arena<ddm::DsConst> dconst;
    Arena<ddm::GenConst> gconst;
    Arena<int32_t> steps, group;
    Arena<double> bound, dbg_z, export_buf, ev_scratch, ev_means, ev_ds_stats, ev_pairs;
    Arena<float> ev_path;
    Arena<int64_t> dbg_off;
    Arena<uint32_t> philox_buf;
    Arena<unsigned long long> hist;
    unsigned long long *counters = nullptr;       // device: [0] work counter, [1..] stats
    unsigned long long *counters_host = nullptr;  // pinned mirror

    // uploaded parameters
    int model = -1;
    int64_t n_datasets = 0;
    int n_params = 0;
    bool have_params = false;
    bool degenerate_noise = false;  // some dataset has dc == 0 (or denormal / non-finite): no unit-scaled state

    // shared-increment mode
    bool dbg_on = false;
    size_t dbg_n = 0;
    int64_t dbg_trials = 0;

    // last run
    void *out = nullptr;
    size_t out_cap = 0, out_bytes = 0;
    bool have_run = false, out64 = true, have_steps = false, stats_pending = false;
    int64_t run_rows = 0, run_datasets = 0, run_trials = 0;  // rows = trials in total
    int run_model = -1;                                       // ddm_model of the last run (kRunModelEvidence for evidence runs)
    int64_t run_cols = 2;                                     // values per trial (2 + n_obs for evidence runs)
    bool run_trialwise = false;
    ddm_stats stats{};

    bool out_resident = false;  // false after a pipelined host-destined run

    // chunked compute / copy pipeline of ddm_simulate (large host-destined batches)
    cudaStream_t copy_stream = nullptr;
    cudaStream_t pipe_stream2 = nullptr;           // odd chunks: their kernels fill the SMs the even chunk's tail frees
    unsigned long long *work_counter2 = nullptr;   // and claim work from their own counter
    cudaEvent_t pipe_ready = nullptr;
    void *pipe_buf[2] = {nullptr, nullptr};
    size_t pipe_cap[2] = {0, 0};
    cudaEvent_t pipe_kernel_done[2] = {nullptr, nullptr}, pipe_copy_done[3] = {nullptr, nullptr, nullptr};
    // compact wire (ddm_wire.cuh): pinned staging for three chunks in flight and the host decode threads
    void *wire_host[3] = {nullptr, nullptr, nullptr};
    size_t wire_cap[3] = {0, 0, 0};
    ddm::HostWorkers *workers = nullptr;
    int tune_host_decode = 0;  // 0 automatic thread count, > 0 that many threads, < 0 plain 16-byte rows over PCIe
    double *train_stage = nullptr;  // pinned: ddm_training_batch's prior draws on their way to the caller's array
    size_t train_stage_cap = 0;
    cudaStream_t hist_stream = nullptr;  // streamed ddm_simulate_histogram: each chunk's reduction runs beside the next chunk's kernel

    // tuning (0 = automatic)
    int tune_threshold = 0, tune_blocks_per_sm = 0, tune_tile = 0;
    // -1 = automatic (tile kernel; latency kernel for launches of at most kLatencyMaxRows trials), 0 = tile kernel,
    // 1 = round-1 persistent kernel (A/B measurements), 2 = latency kernel at any size
    int tune_kernel_variant = -1;
    bool trialwise_degenerate = false;  // last trialwise call: some group has dc == 0 (no noise unit)
    int64_t tune_pipeline_min_rows = -1, tune_pipeline_chunk_rows = -1;  // < 0: default
};

namespace {

// Refill threshold: the emit + refill pass (~60 issue slots) is worth taking once `thr` lanes idle.  With lanes
// finishing at rate r per six-step block the cost per block is ~1.6 (thr - 1) idle-lane slots + 60 r / thr pass
// slots, minimal near thr = sqrt(37.5 r) = 85 / sqrt(steps per trial).  Under the reference's priors a trial
// takes ~0.27 / dt steps (28 at dt = .01, 258 at dt = .001), hence 164 sqrt(dt).  Measured optima on B200: 5-6 at
// dt = .001, 12-16 at dt = .01.  (An in-kernel estimate of r was tried: 2 % slower on the sweep.)  Results never
// depend on the threshold; ddm_set_tuning overrides it.
int default_refill_threshold(double dt, bool legacy = false) {
    if (legacy) {
        const int thr = (int)std::lround(164.0 * std::sqrt(dt));
        return thr < 2 ? 2 : (thr > 16 ? 16 : thr);
    }
    // tile kernel: finished lanes are refilled in place (~45 issue slots per refill, no loop re-entry), so the optimum sits
    // lower than the round-1 kernel's: measured 3-4 at dt = .001 and 8 at dt = .01 for every model family
    // (profiles/r02_ab_kernels.jsonl); 42 dt^0.36 passes through both
    const int thr = (int)std::lround(42.0 * std::pow(dt, 0.36));
    return thr < 2 ? 2 : (thr > 10 ? 10 : thr);
}

int fail(ddm_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define DDM_CUDA(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? DDM_ERR_NOMEM : DDM_ERR_CUDA,      \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int ensure_output(ddm_ctx *ctx, size_t bytes) {
    // reuse the resident buffer only if it is about the right size: a DLPack hand-off gives the whole buffer to the
    // consumer, and a 256 KB training batch must not travel in (and keep alive) the 8 GB buffer of an earlier sweep
    if (ctx->out && ctx->out_cap >= bytes && ctx->out_cap <= 2 * bytes + (1u << 20)) return DDM_OK;
    if (ctx->out) {
        ctx->pool->give(ctx->out, ctx->out_cap);
        ctx->out = nullptr;
        ctx->out_cap = 0;
    }
    size_t cap = 0;
    void *p = ctx->pool->take(bytes, &cap);
    if (!p) {
        cap = bytes < 256 ? 256 : bytes;
        DDM_CUDA(ctx, cudaMalloc(&p, cap));
    }
    ctx->out = p;
    ctx->out_cap = cap;
    return DDM_OK;
}

// Argument checks and the launch-invariant part of RunArgs, shared by every entry point.
int build_args(ddm_ctx *ctx, int model, int64_t n_datasets, int64_t n_trials, double dt, int max_steps, uint64_t seed,
               uint64_t dataset_offset, uint64_t trial_offset, int precision, int flags, ddm::RunArgs &a) {
    if (precision != 32 && precision != 64) return fail(ctx, DDM_ERR_INVALID, "precision must be 32 or 64, got %d", precision);
    if (!(dt > 0.0) || !std::isfinite(dt)) return fail(ctx, DDM_ERR_INVALID, "dt must be positive and finite");
    if (max_steps < 0) return fail(ctx, DDM_ERR_INVALID, "max_steps must be >= 0");
    if (n_trials < 0 || n_datasets < 0) return fail(ctx, DDM_ERR_INVALID, "negative shape");
    const bool trialwise = (model == DDM_MODEL_TRIALWISE);
    const int64_t rows = trialwise ? n_trials : n_datasets * n_trials;
    if (n_trials > 0xffffffffLL || n_datasets > 0xffffffffLL) return fail(ctx, DDM_ERR_INVALID, "shape exceeds 2^32");
    // 64-bit global indices: the low 32 bits are a Philox counter word, bits 32..55 ride in the stream word
    // (PhiloxKey::c3_hi), launch-uniform -- so a launch must not straddle a multiple of 2^32.  For the trialwise
    // model the 64-bit index is the trial's (dataset word = its high part's low 32 bits is not needed: see below).
    uint64_t index_hi = 0;
    if (trialwise) {
        index_hi = trial_offset >> 32;
        trial_offset &= 0xffffffffULL;
        if (dataset_offset != 0) return fail(ctx, DDM_ERR_INVALID, "trialwise runs are keyed by trial_offset only");
    } else {
        index_hi = dataset_offset >> 32;
        dataset_offset &= 0xffffffffULL;
    }
    if (index_hi >= (1ULL << 24)) return fail(ctx, DDM_ERR_INVALID, "global dataset / trial index must stay below 2^56");
    if (dataset_offset + (uint64_t)n_datasets > 0x100000000ULL)
        return fail(ctx, DDM_ERR_INVALID,
                    "a launch may not straddle a multiple of 2^32 datasets (dataset_offset mod 2^32 + n_datasets > 2^32): "
                    "start the batch at the next multiple");
    if (trial_offset + (uint64_t)n_trials > 0x100000000ULL)
        return fail(ctx, DDM_ERR_INVALID, "trial_offset mod 2^32 + n_trials must not exceed 2^32 (Philox counter word)");
    if (ctx->dbg_on && ctx->dbg_trials != rows)
        return fail(ctx, DDM_ERR_INVALID, "shared-increment buffer was set for %lld trials, run has %lld",
                    (long long)ctx->dbg_trials, (long long)rows);
    a = ddm::RunArgs{};
    a.dconst = ctx->dconst.p;
    a.gconst = ctx->gconst.p;
    a.params = ctx->params.p;
    a.group = ctx->group.p;
    a.bound_in = ctx->bound.p;
    a.dbg_z = ctx->dbg_on ? ctx->dbg_z.p : nullptr;
    a.dbg_off = ctx->dbg_on ? ctx->dbg_off.p : nullptr;
    a.dbg_n = ctx->dbg_on ? ctx->dbg_n : 0;
    a.work_counter = ctx->counters;
    a.stats = ctx->counters + 1;
    a.n_datasets = (uint32_t)n_datasets;
    a.n_trials = (uint32_t)n_trials;
    a.n_params = (uint32_t)(trialwise ? 4 : ctx->n_params);
    a.dataset_offset = (uint32_t)dataset_offset;
    a.trial_offset = (uint32_t)trial_offset;
    a.key = ddm::make_philox_key((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)index_hi);
    a.max_steps = (uint32_t)max_steps;
    a.model = model;
    a.flags = flags;
    a.dt = dt;
    a.sqrt_dt = std::sqrt(dt);
    return DDM_OK;
}

// Splits a 64-bit global dataset index into the 32-bit counter word and the launch-uniform high part.
int split_index(ddm_ctx *ctx, uint64_t offset, int64_t count, uint32_t *lo, uint32_t *hi) {
    if ((offset >> 32) >= (1ULL << 24)) return fail(ctx, DDM_ERR_INVALID, "global dataset index must stay below 2^56");
    if ((offset & 0xffffffffULL) + (uint64_t)count > 0x100000000ULL)
        return fail(ctx, DDM_ERR_INVALID,
                    "a launch may not straddle a multiple of 2^32 datasets (dataset_offset mod 2^32 + n_datasets > 2^32): "
                    "start the batch at the next multiple");
    *lo = (uint32_t)offset;
    *hi = (uint32_t)(offset >> 32);
    return DDM_OK;
}

bool uses_dconst(const ddm_ctx *ctx, int model, int precision) {
    if (model == DDM_MODEL_TRIALWISE) return precision == 32 && !ctx->dbg_on && !ctx->trialwise_degenerate;
    return precision == 32 && !ctx->dbg_on;
}

// Enqueue the simulator kernel for the datasets described by `a` (pointers already offset to the
// range) on the ctx stream.  The stats counters accumulate; only the work counter is reset.
constexpr int64_t kLatencyMaxRows = 256 << 10;  // it leads up to here for every model, scripts/r02_latency_probe.py
// (flags: the ddm_flags of the call plus the internal ones already set in RunArgs)
bool takes_latency_kernel(const ddm_ctx *ctx, int model, int64_t rows, int precision, int flags, uint32_t max_steps) {
    const bool trialwise = (model == DDM_MODEL_TRIALWISE);
    const bool degenerate = trialwise ? ctx->trialwise_degenerate : ctx->degenerate_noise;
    const bool persistent = precision == 32 && !ctx->dbg_on && !degenerate && !(flags & (DDM_FLAG_FORCE_GENERIC | DDM_FLAG_OUT_STATE)) &&
                            max_steps <= ddm::TILE_MAX_STEPS;
    return persistent && kind_of(model) != ddm::KIND_GENERAL && !(flags & ddm::FLAG_WIRE_COMPACT) && rows > 0 &&
           (ctx->tune_kernel_variant == 2 || (ctx->tune_kernel_variant == -1 && rows <= kLatencyMaxRows));
}
int launch_sim(ddm_ctx *ctx, ddm::RunArgs &a, int precision, ddm_stats &st, cudaStream_t stream = nullptr) {
    if (!stream) stream = ctx->stream;
    const int model = a.model, flags = a.flags;
    const bool trialwise = (model == DDM_MODEL_TRIALWISE);
    const int kind = kind_of(model);
    const bool out64 = !(flags & DDM_FLAG_OUT_F32);
    const int64_t n_datasets = a.n_datasets, n_trials = a.n_trials;
    const int64_t rows = trialwise ? n_trials : n_datasets * n_trials;
    if (rows == 0) return DDM_OK;
    // OUT_STATE (validation) needs the state at exactly max_steps for timeouts: the generic kernel stops there
    const bool degenerate = trialwise ? ctx->trialwise_degenerate : ctx->degenerate_noise;
    if (degenerate) a.flags |= ddm::FLAG_REFERENCE_ARITHMETIC;
    const bool persistent = precision == 32 && !ctx->dbg_on && !degenerate &&
                            !(flags & (DDM_FLAG_FORCE_GENERIC | DDM_FLAG_OUT_STATE)) && (uint32_t)a.max_steps <= ddm::TILE_MAX_STEPS;
    // Small launches -- one trial per lane or fewer, nothing to refill: the reference's own batch sizes -- last as long as
    // their longest trial's dependent chain; the latency kernel (speculative six-step blocks, one thread per trial) is
    // built for that (scripts/r02_latency_probe.py).  Same bits as the persistent kernels.
    const bool latency = takes_latency_kernel(ctx, model, rows, precision, a.flags, a.max_steps);
    if (latency) {
        DDM_CUDA(ctx, ddm::launch_latency(a, kind, out64, (uint64_t)rows, stream));
        st.used_persistent = 1;  // a production kernel, not the validation twin
        st.scheduler = 3;
        st.grid = (int)(((uint64_t)rows + 127) / 128);
        st.block = 128;
        st.kernel_launches++;
        return DDM_OK;
    }
    if (persistent && kind != ddm::KIND_GENERAL && !a.dconst)
        return fail(ctx, DDM_ERR_STATE, "internal: per-dataset constants missing for a persistent launch");
    DDM_CUDA(ctx, cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), stream));
    if (persistent) {
        const int block = ddm::persistent_block_size();
        // Two schedulers of the same trials, bit-identical results.  The tile kernel (round 2) is the production one for
        // every model family: +3 % on the sweep, +7 % on the basic model at dt = .01, +18-20 % on the per-trial-boundary /
        // per-trial-dc models at dt = .01 (profiles/r02_ab_kernels.jsonl); the round-1 kernel stays selectable for A/B.
        const bool legacy = !trialwise && ctx->tune_kernel_variant == 1;
        const uint64_t warps_needed = ((uint64_t)rows + 31) / 32;
        const uint64_t blocks_needed = (warps_needed + (block / 32) - 1) / (block / 32);
        // tile: consecutive trials of one dataset handed out per atomic claim (and, in the tile kernel, set up and
        // written out together).  Large batches use the largest tile the shared-memory buffers allow (fewest
        // stragglers); small ones shrink it so that every warp of the grid still finds a few tiles.  (Whole-dataset
        // claims measured 15 % slower on the sweep: per-dataset cost varies 400x, coarse claims unbalance the tail.)
        const uint32_t tile_max = legacy ? 0x40000000u : ddm::tile_kernel_max_tile(kind);
        uint32_t tile;
        if (ctx->tune_tile > 0) {
            tile = (uint32_t)ctx->tune_tile;
        } else if (legacy) {
            tile = 64u;
        } else {
            const uint64_t warps_grid = std::min<uint64_t>((uint64_t)ctx->sm_count * 6 * (block / 32), warps_needed);
            const uint64_t want = (uint64_t)rows / (4 * warps_grid);
            tile = want >= tile_max ? tile_max : (want <= 32 ? 32u : (uint32_t)want & ~31u);
        }
        if (tile > tile_max) tile = tile_max;
        if (tile > (uint32_t)n_trials) tile = (uint32_t)n_trials;
        if (tile == 0) tile = 1;
        if (legacy)
            while (((uint64_t)n_trials + tile - 1) / tile * (uint64_t)n_datasets > 0xffffffffULL) tile *= 2;  // 32-bit item index
        a.tile = tile;
        a.tiles_per_dataset = (uint32_t)((n_trials + tile - 1) / tile);
        a.n_items = (uint64_t)a.tiles_per_dataset * (uint64_t)n_datasets;
        if (a.n_items > 0xffffffffULL) return fail(ctx, DDM_ERR_INVALID, "batch too large for one launch (more than 2^32 trial tiles)");
        a.refill_threshold = ctx->tune_threshold > 0 ? ctx->tune_threshold : default_refill_threshold(a.dt, legacy);
        if (a.refill_threshold > 32) a.refill_threshold = 32;
        const size_t smem = legacy ? 0 : ddm::tile_kernel_smem_bytes(kind, block);
        int per_sm = ctx->tune_blocks_per_sm;
        const int max_per_sm = legacy ? ddm::persistent_max_blocks_per_sm(kind, out64, block)
                                      : ddm::tile_kernel_max_blocks_per_sm(kind, out64, block, smem);
        if (max_per_sm <= 0) return fail(ctx, DDM_ERR_CUDA, "occupancy query failed for the persistent kernel");
        if (per_sm <= 0 || per_sm > max_per_sm) per_sm = max_per_sm;
        uint64_t grid = (uint64_t)ctx->sm_count * per_sm;
        if (grid > blocks_needed) grid = blocks_needed;
        if (grid < 1) grid = 1;
        if (legacy) DDM_CUDA(ctx, ddm::launch_persistent(a, kind, out64, (int)grid, block, stream));
        else DDM_CUDA(ctx, ddm::launch_tile(a, kind, out64, (int)grid, block, smem, stream));
        st.used_persistent = 1;
        st.scheduler = legacy ? 1 : 2;
        st.grid = (int)grid;
        st.block = block;
        st.refill_threshold = a.refill_threshold;
        st.tile = (int)tile;
    } else {
        DDM_CUDA(ctx, ddm::launch_generic(a, kind, precision == 64, ctx->dbg_on, out64, (uint64_t)rows, stream));
        st.grid = (int)(((uint64_t)rows + 127) / 128);
        st.block = 128;
    }
    st.kernel_launches++;
    return DDM_OK;
}

// Shared by ddm_run and ddm_simulate_trialwise: one launch over everything, output resident.
int run_common(ddm_ctx *ctx, int model, int64_t n_datasets, int64_t n_trials, double dt, int max_steps,
               uint64_t seed, uint64_t dataset_offset, uint64_t trial_offset, int precision, int flags,
               int n_groups) {
    ddm::RunArgs a;
    int rc = build_args(ctx, model, n_datasets, n_trials, dt, max_steps, seed, dataset_offset, trial_offset, precision, flags, a);
    if (rc) return rc;
    const bool trialwise = (model == DDM_MODEL_TRIALWISE);
    const int64_t rows = trialwise ? n_trials : n_datasets * n_trials;
    const bool out64 = !(flags & DDM_FLAG_OUT_F32);
    const int cols = n_cols_of(model);
    const size_t out_bytes = (size_t)rows * cols * (out64 ? 8 : 4);
    rc = ensure_output(ctx, out_bytes);
    if (rc) return rc;
    const bool keep_steps = (flags & DDM_FLAG_KEEP_STEPS) != 0;
    if (keep_steps) DDM_CUDA(ctx, ctx->steps.reserve((size_t)rows));
    a.out = ctx->out;
    a.steps_out = keep_steps ? ctx->steps.p : nullptr;

    ddm_stats st{};
    st.n_trials = (uint64_t)rows;
    DDM_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT), ctx->stream));
    if (rows > 0) {
        if (takes_latency_kernel(ctx, model, rows, precision, flags, (uint32_t)max_steps)) {
            a.dconst = nullptr;  // a launch this small forms its constants itself: no prep_kernel in front of it (-4 us of ~40)
        } else if (uses_dconst(ctx, model, precision)) {
            if (trialwise) {  // per-participant constants; the boundary comes per trial
                DDM_CUDA(ctx, ctx->dconst.reserve((size_t)(n_groups > 0 ? n_groups : 1)));
                a.dconst = ctx->dconst.p;
                DDM_CUDA(ctx, ddm::launch_prep(ctx->params.p, ctx->dconst.p, (uint32_t)n_groups, 4u, model, dt, ctx->stream));
            } else if (model == DDM_MODEL_GENERAL) {
                DDM_CUDA(ctx, ctx->gconst.reserve((size_t)n_datasets));
                a.gconst = ctx->gconst.p;
                DDM_CUDA(ctx, ddm::launch_prep_general(ctx->params.p, ctx->gconst.p, (uint32_t)n_datasets, dt, ctx->stream));
            } else {
                DDM_CUDA(ctx, ctx->dconst.reserve((size_t)n_datasets));
                a.dconst = ctx->dconst.p;
                DDM_CUDA(ctx, ddm::launch_prep(ctx->params.p, ctx->dconst.p, (uint32_t)n_datasets, (uint32_t)ctx->n_params,
                                               model, dt, ctx->stream));
            }
            st.kernel_launches++;
        }
        DDM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        rc = launch_sim(ctx, a, precision, st);
        if (rc) return rc;
        DDM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    }
    DDM_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, ctx->counters, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats = st;
    ctx->stats_pending = true;
    ctx->have_run = true;
    ctx->out64 = out64;
    ctx->out_bytes = out_bytes;
    ctx->have_steps = keep_steps;
    ctx->run_rows = rows;
    ctx->run_datasets = n_datasets;
    ctx->run_trials = n_trials;
    ctx->run_trialwise = trialwise;
    ctx->run_cols = cols;
    ctx->run_model = model;
    ctx->out_resident = true;
    return DDM_OK;
}

// Large host-destined batches: datasets are simulated in chunks into two device buffers while
// the previous chunk is copied to the host on a second stream, so the batch costs
// max(kernel, PCIe) instead of kernel + PCIe and needs 2 chunks of HBM instead of the batch.
constexpr int64_t kPipelineMinRows = 8ll << 20;     // below this one launch + one copy is as fast
// Compact records pay off from 10^6 trials on for the (rt, choice) layout -- 1024 x 1000 rows 0.44 -> 0.32-0.37 ms in two
// chunks -- and from 4 Mi for the (signed rt, external column) layout, whose 8-byte records and wider decode only break
// even at 10^6 (scripts/r02_c3_pipeline_ab.py, profiles/r02_c3_pipeline_ab.txt); 256 x 1000 gains from nothing.
constexpr int64_t kCompactMinRows = 1000000;
constexpr int64_t kCompactMinRowsExt = 4ll << 20;
// Into a page-locked destination two plain chunks already beat one launch + one copy at 10^6 trials (C3: 0.49 -> 0.44 ms);
// into pageable memory each chunk's copy blocks the host and nothing overlaps, so the old threshold stays.
constexpr int64_t kPinnedPipelineMinRows = 1000000;
constexpr int64_t kPipelineMinChunkRows = 2ll << 20;  // a chunk costs ~0.2 ms of launches and kernel tail
constexpr int64_t kPipelineSmallBatchRows = 4ll << 20;     // batches below this are halved / quartered ...
constexpr int64_t kPipelineSmallBatchChunkRows = 512ll << 10;  // ... down to chunks of 512 Ki trials
constexpr int64_t kPipelineChunkRows = 32ll << 20;  // trials per chunk (512 MB of float64 pairs)

// The host thread pool (compact-wire decode, parameter scan), sized by ddm_set_host_decode.
int ensure_workers(ddm_ctx *ctx) {
    int gpus = 1;
    if (cudaGetDeviceCount(&gpus) != cudaSuccess) gpus = 1;
    const int want = ctx->tune_host_decode > 0 ? ctx->tune_host_decode : ddm::host_workers_default_count(gpus);
    if (ctx->workers && ddm::host_workers_size(ctx->workers) != want) {
        ddm::host_workers_destroy(ctx->workers);
        ctx->workers = nullptr;
    }
    if (!ctx->workers) ctx->workers = ddm::host_workers_create(want);
    return ctx->workers ? DDM_OK : DDM_ERR_NOMEM;
}

bool takes_persistent_kernel(const ddm_ctx *ctx, int model, int precision, int flags) {
    return precision == 32 && !ctx->dbg_on && model != DDM_MODEL_TRIALWISE && !ctx->degenerate_noise &&
           !(flags & (DDM_FLAG_FORCE_GENERIC | DDM_FLAG_OUT_STATE));
}

// The streamed path's chunk schedule: (first dataset, datasets) per chunk, covering [0, n_datasets) in order.
std::vector<std::pair<int64_t, int64_t>> pipeline_chunks(int64_t n_datasets, int64_t n_trials, int64_t chunk_rows) {
    std::vector<std::pair<int64_t, int64_t>> chunks;
    const int64_t per_ds = n_trials > 0 ? n_trials : 1;
    for (int64_t lo = 0; lo < n_datasets;) {
        int64_t rows_c = chunk_rows;
        if (rows_c <= 0) {
            const int64_t quarter = n_datasets * per_ds / 4, half_left = (n_datasets - lo) * per_ds / 2;
            rows_c = rows_c == -2 ? quarter : (rows_c == -3 ? half_left : (quarter < half_left ? quarter : half_left));
            const int64_t min_chunk = n_datasets * per_ds < kPipelineSmallBatchRows ? kPipelineSmallBatchChunkRows : kPipelineMinChunkRows;
            if (rows_c < min_chunk) rows_c = min_chunk;
            if (rows_c > kPipelineChunkRows) rows_c = kPipelineChunkRows;
        }
        int64_t cnt = rows_c / per_ds;
        if (cnt < 1) cnt = 1;
        if (cnt > n_datasets - lo) cnt = n_datasets - lo;
        chunks.emplace_back(lo, cnt);
        lo += cnt;
    }
    return chunks;
}

// Streams, events and the second work counter of the chunked paths (created on first use).
int ensure_pipe_streams(ddm_ctx *ctx) {
    if (ctx->copy_stream) return DDM_OK;
    DDM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    DDM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->pipe_stream2, cudaStreamNonBlocking));
    DDM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->hist_stream, cudaStreamNonBlocking));
    DDM_CUDA(ctx, cudaMalloc(&ctx->work_counter2, sizeof(unsigned long long)));
    DDM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_ready, cudaEventDisableTiming));
    for (int b = 0; b < 2; b++) DDM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_kernel_done[b], cudaEventDisableTiming));
    for (int b = 0; b < 3; b++) DDM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_copy_done[b], cudaEventDisableTiming));
    return DDM_OK;
}

int run_pipelined(ddm_ctx *ctx, const double *params_host, int64_t n_trials, double dt, int max_steps, uint64_t seed,
                  uint64_t dataset_offset, int precision, int flags, void *out_host, bool want_compact) {
    const int model = ctx->model;
    const int64_t n_datasets = ctx->n_datasets;
    ddm::RunArgs base;
    int rc = build_args(ctx, model, n_datasets, n_trials, dt, max_steps, seed, dataset_offset, 0, precision, flags, base);
    if (rc) return rc;
    const bool out64 = !(flags & DDM_FLAG_OUT_F32);
    const int cols = n_cols_of(model);
    const size_t row_bytes = (size_t)cols * (out64 ? 8 : 4);
    // Compact wire (ddm_wire.cuh): two-column models leave the device as (steps, choice[, fp32 draw]) records and
    // host threads write the rows, so the PCIe copy moves 4 or 8 bytes per trial instead of 16.
    const int kind = kind_of(model);
    const bool basic_cols = (kind == ddm::KIND_FIXED || kind == ddm::KIND_DRIFT);
    const bool compact = want_compact && ctx->tune_host_decode >= 0 && cols == 2 && takes_persistent_kernel(ctx, model, precision, flags) &&
                         (uint32_t)max_steps <= ddm::WIRE_MAX_STEPS;
    const size_t wire_bytes = compact ? (basic_cols ? 4 : 8) : row_bytes;
    // Chunk schedule (first dataset, datasets).  A fixed chunk_rows if the caller set one; otherwise each chunk is
    // the smaller of a quarter of the batch and half of what is left, within 2 Mi .. 32 Mi trials: large batches
    // run in 32 Mi chunks (fewest launches and kernel tails), the last chunks shrink so that little copy + decode is
    // left exposed after the last kernel, and mid-size batches overlap kernel, copy and host decode as well
    // (profiles/r01_v9c_midsize_ab.txt, r01_v9d_chunk_policy_ab.txt; codes -2 / -3 select either rule alone).
    const std::vector<std::pair<int64_t, int64_t>> chunks = pipeline_chunks(n_datasets, n_trials, ctx->tune_pipeline_chunk_rows);
    const int64_t n_chunks = (int64_t)chunks.size();
    int64_t chunk_ds = 0;  // the largest chunk sizes the buffers
    for (const auto &c : chunks) chunk_ds = c.second > chunk_ds ? c.second : chunk_ds;
    const size_t chunk_bytes = (size_t)chunk_ds * (size_t)n_trials * wire_bytes;
    // the resident-output buffer is not used by this path: hand it back so the pool can reuse it
    if (ctx->out) {
        ctx->pool->give(ctx->out, ctx->out_cap);
        ctx->out = nullptr;
        ctx->out_cap = 0;
    }
    for (int b = 0; b < 2; b++) {
        if (ctx->pipe_buf[b] && ctx->pipe_cap[b] >= chunk_bytes) continue;
        if (ctx->pipe_buf[b]) cudaFree(ctx->pipe_buf[b]);
        ctx->pipe_buf[b] = nullptr;
        ctx->pipe_cap[b] = 0;
        DDM_CUDA(ctx, cudaMalloc(&ctx->pipe_buf[b], chunk_bytes));
        ctx->pipe_cap[b] = chunk_bytes;
    }
    if (compact) {
        for (int b = 0; b < 3 && b < n_chunks; b++) {
            if (ctx->wire_host[b] && ctx->wire_cap[b] >= chunk_bytes) continue;
            if (ctx->wire_host[b]) cudaFreeHost(ctx->wire_host[b]);
            ctx->wire_host[b] = nullptr;
            ctx->wire_cap[b] = 0;
            DDM_CUDA(ctx, cudaHostAlloc(&ctx->wire_host[b], chunk_bytes, cudaHostAllocDefault));
            ctx->wire_cap[b] = chunk_bytes;
        }
        rc = ensure_workers(ctx);
        if (rc) return rc;
    }
    rc = ensure_pipe_streams(ctx);
    if (rc) return rc;
    ddm_stats st{};
    st.n_trials = (uint64_t)(n_datasets * n_trials);
    DDM_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT), ctx->stream));
    const bool dconst = uses_dconst(ctx, model, precision);
    if (dconst) {
        if (model == DDM_MODEL_GENERAL) DDM_CUDA(ctx, ctx->gconst.reserve((size_t)n_datasets));
        else DDM_CUDA(ctx, ctx->dconst.reserve((size_t)n_datasets));
    }
    DDM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    // Even chunks run on the ctx stream, odd chunks on a second one with their own work counter: a chunk's kernel
    // ends with a tail in which a few warps finish their longest trials (up to max_steps steps, ~0.5 ms at
    // dt = .001) while most SMs idle, and the next chunk's blocks move in as this one's retire.
    DDM_CUDA(ctx, cudaEventRecord(ctx->pipe_ready, ctx->stream));
    DDM_CUDA(ctx, cudaStreamWaitEvent(ctx->pipe_stream2, ctx->pipe_ready, 0));  // the counters are zeroed
    // chunk i: device buffer and stream i % 2, copy-done event (and wire staging buffer) i % 3
    auto enqueue = [&](int64_t i) -> int {
        const int b = (int)(i & 1), s3 = (int)(i % 3);
        cudaStream_t ks = b ? ctx->pipe_stream2 : ctx->stream;
        const int64_t lo = chunks[i].first, cnt = chunks[i].second;
        // device buffer free again?  (compact: the host has already waited for that copy)
        if (!compact && i >= 2) DDM_CUDA(ctx, cudaStreamWaitEvent(ks, ctx->pipe_copy_done[(i - 2) % 3], 0));
        // the chunk's parameters travel with it (a 1e6-dataset batch is 40 MB of pageable host memory: ~4 ms that
        // only the first chunk would otherwise wait for), then its per-dataset constants.  (A copy from pageable
        // memory first waits for this stream's earlier work, i.e. for chunk i - 2: the other stream still holds
        // chunk i - 1, so the GPU stays busy.)
        const size_t p_lo = (size_t)lo * ctx->n_params;
        DDM_CUDA(ctx, cudaMemcpyAsync(ctx->params.p + p_lo, params_host + p_lo, (size_t)cnt * ctx->n_params * sizeof(double),
                                      cudaMemcpyHostToDevice, ks));
        if (dconst) {
            if (model == DDM_MODEL_GENERAL)
                DDM_CUDA(ctx, ddm::launch_prep_general(ctx->params.p + p_lo, ctx->gconst.p + lo, (uint32_t)cnt, dt, ks));
            else
                DDM_CUDA(ctx, ddm::launch_prep(ctx->params.p + p_lo, ctx->dconst.p + lo, (uint32_t)cnt, (uint32_t)ctx->n_params, model, dt, ks));
            st.kernel_launches++;
        }
        ddm::RunArgs a = base;
        if (b) a.work_counter = ctx->work_counter2;
        a.params = ctx->params.p + (size_t)lo * ctx->n_params;
        a.dconst = (dconst && model != DDM_MODEL_GENERAL) ? ctx->dconst.p + lo : nullptr;
        a.gconst = (dconst && model == DDM_MODEL_GENERAL) ? ctx->gconst.p + lo : nullptr;
        a.n_datasets = (uint32_t)cnt;
        a.dataset_offset = (uint32_t)(dataset_offset + (uint64_t)lo);
        a.out = ctx->pipe_buf[b];
        a.steps_out = nullptr;
        if (compact) a.flags |= ddm::FLAG_WIRE_COMPACT;
        const int rc2 = launch_sim(ctx, a, precision, st, ks);
        if (rc2) return rc2;
        DDM_CUDA(ctx, cudaEventRecord(ctx->pipe_kernel_done[b], ks));
        DDM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->pipe_kernel_done[b], 0));
        void *dst = compact ? ctx->wire_host[s3] : static_cast<void *>(static_cast<char *>(out_host) + (size_t)lo * (size_t)n_trials * row_bytes);
        DDM_CUDA(ctx, cudaMemcpyAsync(dst, ctx->pipe_buf[b], (size_t)cnt * (size_t)n_trials * wire_bytes, cudaMemcpyDeviceToHost,
                                      ctx->copy_stream));
        DDM_CUDA(ctx, cudaEventRecord(ctx->pipe_copy_done[s3], ctx->copy_stream));
        return DDM_OK;
    };
    if (!compact) {
        for (int64_t i = 0; i < n_chunks; i++) {
            rc = enqueue(i);
            if (rc) return rc;
        }
    } else {
        // two chunks stay queued on the GPU while the host threads write the rows of the chunk that has landed
        struct HotWorkers {
            ddm::HostWorkers *w;
            explicit HotWorkers(ddm::HostWorkers *w_) : w(w_) { ddm::host_workers_begin(w); }
            ~HotWorkers() { ddm::host_workers_end(w); }
        } hot(ctx->workers);
        for (int64_t i = 0; i < 2 && i < n_chunks; i++) {
            rc = enqueue(i);
            if (rc) return rc;
        }
        for (int64_t i = 0; i < n_chunks; i++) {
            DDM_CUDA(ctx, cudaEventSynchronize(ctx->pipe_copy_done[i % 3]));
            if (i + 2 < n_chunks) {
                rc = enqueue(i + 2);
                if (rc) return rc;
            }
            const int64_t lo = chunks[i].first;
            ddm::WireDecode job;
            job.wire = ctx->wire_host[i % 3];
            job.out = static_cast<char *>(out_host) + (size_t)lo * (size_t)n_trials * row_bytes;
            job.params = params_host + (size_t)lo * ctx->n_params;
            job.n_params = ctx->n_params;
            job.tau_col = 3;
            job.n_datasets = chunks[i].second;
            job.n_trials = n_trials;
            job.dt = dt;
            job.basic = basic_cols;
            job.out64 = out64;
            job.timeout_choice_one = (flags & DDM_FLAG_TIMEOUT_CHOICE_ONE) != 0;
            ddm::wire_decode(ctx->workers, job);
        }
    }
    // the caller's stream sees kernels (both streams) and copies as done: join the copy stream back, then block
    for (int64_t i = (n_chunks > 3 ? n_chunks - 3 : 0); i < n_chunks; i++)
        DDM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_copy_done[i % 3], 0));
    DDM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, ctx->counters, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    st.d2h_bytes = (uint64_t)n_datasets * (uint64_t)n_trials * wire_bytes;
    st.host_decode_threads = compact ? ddm::host_workers_size(ctx->workers) : 0;
    ctx->stats = st;
    ctx->stats_pending = true;
    ctx->have_run = true;
    ctx->out64 = out64;
    ctx->out_bytes = 0;
    ctx->have_steps = false;
    ctx->run_rows = n_datasets * n_trials;
    ctx->run_datasets = n_datasets;
    ctx->run_trials = n_trials;
    ctx->run_trialwise = false;
    ctx->run_cols = cols;
    ctx->run_model = model;
    ctx->out_resident = false;  // the batch went to the host chunk by chunk
    return DDM_OK;
}

// ddm_simulate_histogram for large batches (C5 as SURVEY.md section 8d specifies it, end to end): the batch stays
// resident as one buffer, exactly as after ddm_run, but it is produced in a few chunks of datasets so that nothing but
// the first chunk's parameters has to cross PCIe before the GPU starts and nothing but the last chunk's reduction is
// left when it stops.  Chunk i: parameters H2D on the copy stream -> prep + simulator kernel on the ctx stream (even i)
// or the second kernel stream (odd i; own work counter, so its blocks move in as the previous chunk's retire) ->
// rt_histogram_kernel over the chunk's rows on the histogram stream, accumulating into one device histogram (it runs
// in the SM time the next kernel's start and the previous kernel's tail leave).  Schedule: a sixteenth of the batch
// first (12 ms of kernel at 1e9 trials: time to upload the next chunk's 19 MB from pageable memory), then half of what
// is left each time down to 32 Mi trials (six chunks at 1e9 trials; the last reduction covers 6 % of the rows, 0.1 ms).
// Measured at 1e6 x 1000 (scripts/r02_hist_stream_ab.py): 199.0-199.8 ms per call with one upload, one launch and one
// reduction, 197.4 ms chunked.  Results do not depend on the chunking: a trial's Philox counters are its global
// (dataset, trial) indices and the histogram is additive.
constexpr int64_t kHistStreamMinRows = 64ll << 20;
constexpr int64_t kHistStreamMinChunkRows = 32ll << 20;  // a chunk boundary costs ~0.2 ms (profiles/r02_hist_stream_ab.txt)

std::vector<std::pair<int64_t, int64_t>> histogram_chunks(int64_t n_datasets, int64_t n_trials, int64_t min_chunk_rows) {
    std::vector<std::pair<int64_t, int64_t>> chunks;
    const int64_t per_ds = n_trials > 0 ? n_trials : 1;
    if (min_chunk_rows <= 0) min_chunk_rows = kHistStreamMinChunkRows;
    const int64_t min_cnt = std::max<int64_t>(1, (min_chunk_rows + per_ds - 1) / per_ds);
    for (int64_t lo = 0; lo < n_datasets;) {
        const int64_t left = n_datasets - lo;
        int64_t cnt = chunks.empty() ? n_datasets / 16 : left / 2;
        if (cnt < min_cnt) cnt = min_cnt;
        if (left - cnt < min_cnt) cnt = left;  // no crumbs
        chunks.emplace_back(lo, cnt);
        lo += cnt;
    }
    return chunks;
}

int run_histogram_streamed(ddm_ctx *ctx, const double *params_host, int64_t n_trials, double dt, int max_steps, uint64_t seed,
                           uint64_t dataset_offset, int precision, int flags, int n_bins, double rt_max, uint64_t *hist_host) {
    const int model = ctx->model;
    const int64_t n_datasets = ctx->n_datasets;
    ddm::RunArgs base;
    int rc = build_args(ctx, model, n_datasets, n_trials, dt, max_steps, seed, dataset_offset, 0, precision, flags, base);
    if (rc) return rc;
    const bool out64 = !(flags & DDM_FLAG_OUT_F32);
    const int cols = n_cols_of(model);
    const size_t row_bytes = (size_t)cols * (out64 ? 8 : 4);
    const int64_t rows = n_datasets * n_trials;
    rc = ensure_output(ctx, (size_t)rows * row_bytes);
    if (rc) return rc;
    rc = ensure_pipe_streams(ctx);
    if (rc) return rc;
    const size_t n_cells = 2 * (size_t)n_bins + 2;
    DDM_CUDA(ctx, ctx->hist.reserve(n_cells));
    const bool dconst = uses_dconst(ctx, model, precision);
    if (dconst) DDM_CUDA(ctx, ctx->dconst.reserve((size_t)n_datasets));
    const bool basic = model == DDM_MODEL_BASIC || model == DDM_MODEL_ETA;
    const std::vector<std::pair<int64_t, int64_t>> chunks = histogram_chunks(n_datasets, n_trials, ctx->tune_pipeline_chunk_rows);
    const int64_t n_chunks = (int64_t)chunks.size();
    // one event per chunk and stage; they live for this call only
    struct Events {
        std::vector<cudaEvent_t> v;
        ~Events() { for (cudaEvent_t e : v) cudaEventDestroy(e); }
    } up, kd;
    for (int64_t i = 0; i < n_chunks; i++) {
        cudaEvent_t e = nullptr;
        DDM_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        up.v.push_back(e);
        DDM_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        kd.v.push_back(e);
    }
    ddm_stats st{};
    st.n_trials = (uint64_t)rows;
    DDM_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT), ctx->stream));
    DDM_CUDA(ctx, cudaMemsetAsync(ctx->hist.p, 0, n_cells * sizeof(unsigned long long), ctx->stream));
    DDM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    // the side streams start behind whatever the caller's stream still holds (earlier runs read the same arenas)
    DDM_CUDA(ctx, cudaEventRecord(ctx->pipe_ready, ctx->stream));
    DDM_CUDA(ctx, cudaStreamWaitEvent(ctx->pipe_stream2, ctx->pipe_ready, 0));
    DDM_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->pipe_ready, 0));
    DDM_CUDA(ctx, cudaStreamWaitEvent(ctx->hist_stream, ctx->pipe_ready, 0));
    for (int64_t i = 0; i < n_chunks; i++) {
        const int b = (int)(i & 1);
        cudaStream_t ks = b ? ctx->pipe_stream2 : ctx->stream;
        const int64_t lo = chunks[i].first, cnt = chunks[i].second;
        const size_t p_lo = (size_t)lo * ctx->n_params;
        DDM_CUDA(ctx, cudaMemcpyAsync(ctx->params.p + p_lo, params_host + p_lo, (size_t)cnt * ctx->n_params * sizeof(double),
                                      cudaMemcpyHostToDevice, ctx->copy_stream));
        DDM_CUDA(ctx, cudaEventRecord(up.v[i], ctx->copy_stream));
        DDM_CUDA(ctx, cudaStreamWaitEvent(ks, up.v[i], 0));
        if (dconst) {
            DDM_CUDA(ctx, ddm::launch_prep(ctx->params.p + p_lo, ctx->dconst.p + lo, (uint32_t)cnt, (uint32_t)ctx->n_params, model, dt, ks));
            st.kernel_launches++;
        }
        ddm::RunArgs a = base;
        if (b) a.work_counter = ctx->work_counter2;
        a.params = ctx->params.p + p_lo;
        a.dconst = dconst ? ctx->dconst.p + lo : nullptr;
        a.n_datasets = (uint32_t)cnt;
        a.dataset_offset = (uint32_t)(dataset_offset + (uint64_t)lo);
        char *chunk_out = static_cast<char *>(ctx->out) + (size_t)lo * (size_t)n_trials * row_bytes;
        a.out = chunk_out;
        a.steps_out = nullptr;
        rc = launch_sim(ctx, a, precision, st, ks);
        if (rc) return rc;
        DDM_CUDA(ctx, cudaEventRecord(kd.v[i], ks));
        DDM_CUDA(ctx, cudaStreamWaitEvent(ctx->hist_stream, kd.v[i], 0));
        DDM_CUDA(ctx, ddm::launch_rt_histogram(chunk_out, out64, (uint64_t)cnt * (uint64_t)n_trials, (uint32_t)cols, basic, (uint32_t)n_bins,
                                               rt_max, ctx->hist.p, ctx->sm_count, ctx->hist_stream));
        st.kernel_launches++;
    }
    // the histogram stream has waited for every kernel: join it back into the caller's stream
    DDM_CUDA(ctx, cudaEventRecord(ctx->pipe_kernel_done[0], ctx->hist_stream));
    DDM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_kernel_done[0], 0));
    DDM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(hist_host, ctx->hist.p, n_cells * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, ctx->counters, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats = st;
    ctx->stats_pending = true;
    ctx->have_run = true;
    ctx->out64 = out64;
    ctx->out_bytes = (size_t)rows * row_bytes;
    ctx->have_steps = false;
    ctx->run_rows = rows;
    ctx->run_datasets = n_datasets;
    ctx->run_trials = n_trials;
    ctx->run_trialwise = false;
    ctx->run_cols = cols;
    ctx->run_model = model;
    ctx->out_resident = true;
    return DDM_OK;
}

int finish_stats(ddm_ctx *ctx) {
    if (!ctx->stats_pending) return DDM_OK;
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned long long *c = ctx->counters_host + 1;
    ctx->stats.total_steps = c[ddm::STAT_STEPS];
    ctx->stats.n_timeouts = c[ddm::STAT_TIMEOUTS];
    ctx->stats.n_upper = c[ddm::STAT_UPPER];
    ctx->stats.reject_cap_hits = c[ddm::STAT_REJECT_CAP];
    ctx->stats.debug_overruns = c[ddm::STAT_DBG_OVERRUN];
    float ms = 0.f;
    if (ctx->run_rows > 0) DDM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.kernel_ms = ms;
    ctx->stats_pending = false;
    return DDM_OK;
}

}  // namespace

// ---- lifecycle --------------------------------------------------------------------------
DDM_API int ddm_version(void) { return DDM_B200_VERSION; }

DDM_API int ddm_philox_rounds(void) { return DDM_PHILOX_ROUNDS; }

DDM_API const char *ddm_last_error(const ddm_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

DDM_API int ddm_create(int device, ddm_ctx **out) {
    if (!out) return fail(nullptr, DDM_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, DDM_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(nullptr, DDM_ERR_INVALID, "device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, DDM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, DDM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    ddm_ctx *ctx = new (std::nothrow) ddm_ctx();
    if (!ctx) return fail(nullptr, DDM_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    DeviceGuard g(device);
    auto bail = [&](const char *what, cudaError_t err) {
        int rc = fail(nullptr, DDM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
        ddm_destroy(ctx);
        return rc;
    };
    ctx->pool = new BufferPool();
    ctx->pool->device = device;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    ctx->stream = ctx->own_stream;
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->counters, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMallocHost(&ctx->counters_host, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT))) != cudaSuccess)
        return bail("cudaMallocHost", e);
    *out = ctx;
    return DDM_OK;
}

DDM_API int ddm_destroy(ddm_ctx *ctx) {
    if (!ctx) return DDM_OK;
    {
        DeviceGuard g(ctx->device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        ctx->params.free_();
        ctx->dconst.free_();
        ctx->gconst.free_();
        ctx->steps.free_();
        ctx->group.free_();
        ctx->bound.free_();
        ctx->dbg_z.free_();
        ctx->export_buf.free_();
        ctx->ev_scratch.free_();
        ctx->ev_means.free_();
        ctx->ev_ds_stats.free_();
        ctx->ev_pairs.free_();
        ctx->ev_path.free_();
        ctx->dbg_off.free_();
        ctx->philox_buf.free_();
        ctx->hist.free_();
        for (int b = 0; b < 2; b++) {
            if (ctx->pipe_buf[b]) cudaFree(ctx->pipe_buf[b]);
            if (ctx->pipe_kernel_done[b]) cudaEventDestroy(ctx->pipe_kernel_done[b]);
        }
        for (int b = 0; b < 3; b++) {
            if (ctx->pipe_copy_done[b]) cudaEventDestroy(ctx->pipe_copy_done[b]);
            if (ctx->wire_host[b]) cudaFreeHost(ctx->wire_host[b]);
        }
        if (ctx->workers) ddm::host_workers_destroy(ctx->workers);
        if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
        if (ctx->pipe_stream2) cudaStreamDestroy(ctx->pipe_stream2);
        if (ctx->hist_stream) cudaStreamDestroy(ctx->hist_stream);
        if (ctx->train_stage) cudaFreeHost(ctx->train_stage);
        if (ctx->work_counter2) cudaFree(ctx->work_counter2);
        if (ctx->pipe_ready) cudaEventDestroy(ctx->pipe_ready);
        if (ctx->counters) cudaFree(ctx->counters);
        if (ctx->counters_host) cudaFreeHost(ctx->counters_host);
        if (ctx->out && ctx->pool) ctx->pool->give(ctx->out, ctx->out_cap);
        if (ctx->ev0) cudaEventDestroy(ctx->ev0);
        if (ctx->ev1) cudaEventDestroy(ctx->ev1);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
        if (ctx->pool) ctx->pool->release();
    }
    delete ctx;
    return DDM_OK;
}

DDM_API int ddm_set_stream(ddm_ctx *ctx, void *cuda_stream) {
    if (!ctx) return DDM_ERR_INVALID;
    DeviceGuard g(ctx->device);
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return DDM_OK;
}

DDM_API int ddm_synchronize(ddm_ctx *ctx) {
    if (!ctx) return DDM_ERR_INVALID;
    DeviceGuard g(ctx->device);
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DDM_OK;
}

DDM_API int ddm_set_tuning(ddm_ctx *ctx, int refill_threshold, int blocks_per_sm, int tile) {
    if (!ctx) return DDM_ERR_INVALID;
    if (refill_threshold < 0 || refill_threshold > 32 || blocks_per_sm < 0 || tile < 0)
        return fail(ctx, DDM_ERR_INVALID, "tuning values out of range");
    ctx->tune_threshold = refill_threshold;
    ctx->tune_blocks_per_sm = blocks_per_sm;
    ctx->tune_tile = tile;
    return DDM_OK;
}

DDM_API int ddm_set_kernel_variant(ddm_ctx *ctx, int variant) {
    if (!ctx) return DDM_ERR_INVALID;
    if (variant < -1 || variant > 2)
        return fail(ctx, DDM_ERR_INVALID,
                    "kernel variant must be -1 (automatic), 0 (tile kernel), 1 (round-1 persistent kernel) or 2 (latency kernel)");
    ctx->tune_kernel_variant = variant;
    return DDM_OK;
}

DDM_API int ddm_set_pipeline(ddm_ctx *ctx, int64_t min_rows, int64_t chunk_rows) {
    if (!ctx) return DDM_ERR_INVALID;
    ctx->tune_pipeline_min_rows = min_rows;
    ctx->tune_pipeline_chunk_rows = chunk_rows;
    return DDM_OK;
}

DDM_API int ddm_set_host_decode(ddm_ctx *ctx, int n_threads) {
    if (!ctx) return DDM_ERR_INVALID;
    if (n_threads > 256) return fail(ctx, DDM_ERR_INVALID, "at most 256 host decode threads, got %d", n_threads);
    ctx->tune_host_decode = n_threads;
    return DDM_OK;
}

// ---- the hot path -----------------------------------------------------------------------
// Validates and registers a parameter batch; copies it to the device unless the caller streams it chunk by chunk.
static int upload_params_impl(ddm_ctx *ctx, int model, const double *params, int64_t n_datasets, int n_params, bool copy) {
    if (!ctx) return DDM_ERR_INVALID;
    const int want = n_params_of(model);
    if (want < 0 || model == DDM_MODEL_TRIALWISE)
        return fail(ctx, DDM_ERR_INVALID, "model %d is not a dataset-wise model (use ddm_simulate_trialwise for 5)", model);
    if (n_params != want) return fail(ctx, DDM_ERR_INVALID, "model %d takes %d parameters per dataset, got %d", model, want, n_params);
    if (n_datasets < 0) return fail(ctx, DDM_ERR_INVALID, "n_datasets < 0");
    if (n_datasets > 0 && !params) return fail(ctx, DDM_ERR_INVALID, "params is NULL");
    DeviceGuard g(ctx->device);
    const size_t n = (size_t)n_datasets * n_params;
    DDM_CUDA(ctx, ctx->params.reserve(n ? n : 1));
    if (n && copy) DDM_CUDA(ctx, cudaMemcpyAsync(ctx->params.p, params, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // The production kernel measures the state in units of sqrt(dt)*dc*sqrt(2 ln 2); a dataset without
    // noise (dc == 0; the reference then runs a deterministic drift) has no such unit and goes through
    // the kernel that keeps the reference's formulas.  (Model 2 draws its per-trial dc > 0 itself.)
    ctx->degenerate_noise = false;
    const int dc_col = (model == DDM_MODEL_BASIC || model == DDM_MODEL_GENERAL) ? 4 : (model == DDM_MODEL_ALPHA_DC ? -1 : 5);
    if (dc_col >= 0) {
        struct Scan {
            const double *params;
            int64_t n_datasets;
            int n_params, dc_col;
            bool general;
            std::atomic<bool> found;
        } scan{params, n_datasets, n_params, dc_col, model == DDM_MODEL_GENERAL, {false}};
        auto slice = [](const void *arg, int id, int n) {
            Scan &sc = *const_cast<Scan *>(static_cast<const Scan *>(arg));
            const int64_t per = (sc.n_datasets + n - 1) / n, lo = per * id, hi = std::min<int64_t>(sc.n_datasets, lo + per);
            for (int64_t d = lo; d < hi; d++) {
                const double dc = sc.params[(size_t)d * sc.n_params + sc.dc_col];
                if (sc.general && sc.params[(size_t)d * sc.n_params + 5] != 0.0) continue;  // redrawn until > 0
                if (!(dc > 1e-30) || !std::isfinite(dc)) {
                    sc.found.store(true, std::memory_order_relaxed);
                    break;
                }
            }
        };
        // a 1e6-dataset batch is 40 MB to walk: 4 ms on one thread, so large batches borrow the decode threads
        if (n_datasets >= (256 << 10) && ctx->tune_host_decode >= 0) {
            int rc = ensure_workers(ctx);
            if (rc) return rc;
            ddm::host_workers_run(ctx->workers, slice, &scan);
        } else {
            slice(&scan, 0, 1);
        }
        ctx->degenerate_noise = scan.found.load();
    }
    ctx->model = model;
    ctx->n_datasets = n_datasets;
    ctx->n_params = n_params;
    ctx->have_params = true;
    return DDM_OK;
}

DDM_API int ddm_upload_params(ddm_ctx *ctx, int model, const double *params, int64_t n_datasets, int n_params) {
    return upload_params_impl(ctx, model, params, n_datasets, n_params, true);
}

DDM_API int ddm_draw_prior(ddm_ctx *ctx, int prior, int64_t n_draws, uint64_t seed, uint64_t draw_offset, double *params_host) {
    if (!ctx) return DDM_ERR_INVALID;
    int n_params, model;
    switch (prior) {
    case DDM_PRIOR_BASIC: case DDM_PRIOR_SWEEP: n_params = 5; model = DDM_MODEL_BASIC; break;
    case DDM_PRIOR_ALPHA: case DDM_PRIOR_ALPHA_DC: case DDM_PRIOR_ALPHA_SCALE2: n_params = 7; model = prior; break;
    case DDM_PRIOR_ALPHA_SCALE: n_params = 8; model = prior; break;
    case DDM_PRIOR_ETA: n_params = 6; model = DDM_MODEL_ETA; break;
    case DDM_PRIOR_EVIDENCE: n_params = 6; model = -1; break;
    default: return fail(ctx, DDM_ERR_INVALID, "unknown prior %d", prior);
    }
    if (n_draws < 0) return fail(ctx, DDM_ERR_INVALID, "n_draws < 0");
    DeviceGuard g(ctx->device);
    const size_t n = (size_t)n_draws * n_params;
    DDM_CUDA(ctx, ctx->params.reserve(n ? n : 1));
    const ddm::PhiloxKey key = ddm::make_philox_key((uint32_t)seed, (uint32_t)(seed >> 32));
    DDM_CUDA(ctx, ddm::launch_prior(ctx->params.p, prior, (uint32_t)n_params, (uint64_t)n_draws, draw_offset, key, ctx->stream));
    ctx->degenerate_noise = false;  // every prior family has dc > 0
    if (model >= 0) {
        ctx->model = model;
        ctx->n_datasets = n_draws;
        ctx->n_params = n_params;
        ctx->have_params = true;
    } else {
        ctx->have_params = false;
    }
    if (params_host && n) {
        DDM_CUDA(ctx, cudaMemcpyAsync(params_host, ctx->params.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return DDM_OK;
}

DDM_API int ddm_run(ddm_ctx *ctx, int64_t n_trials, double dt, int max_steps, uint64_t seed, uint64_t dataset_offset,
                    int precision, int flags) {
    if (!ctx) return DDM_ERR_INVALID;
    if (!ctx->have_params) return fail(ctx, DDM_ERR_STATE, "ddm_run before ddm_upload_params");
    DeviceGuard g(ctx->device);
    return run_common(ctx, ctx->model, ctx->n_datasets, n_trials, dt, max_steps, seed, dataset_offset, 0, precision, flags, 0);
}

DDM_API int ddm_download(ddm_ctx *ctx, void *out_host) {
    if (!ctx) return DDM_ERR_INVALID;
    if (!ctx->have_run || !ctx->out || !ctx->out_resident)
        return fail(ctx, DDM_ERR_STATE, "no output to download (run first; DLPack hand-off moves it away)");
    if (!out_host && ctx->out_bytes) return fail(ctx, DDM_ERR_INVALID, "out_host is NULL");
    DeviceGuard g(ctx->device);
    if (ctx->out_bytes) DDM_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->out, ctx->out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DDM_OK;
}

DDM_API int ddm_simulate(ddm_ctx *ctx, int model, const double *params, int64_t n_datasets, int n_params,
                         int64_t n_trials, double dt, int max_steps, uint64_t seed, uint64_t dataset_offset,
                         int precision, int flags, void *out_host) {
    int rc = upload_params_impl(ctx, model, params, n_datasets, n_params, false);
    if (rc) return rc;
    const bool compact_ok = ctx->tune_host_decode >= 0 && n_cols_of(model) == 2 && takes_persistent_kernel(ctx, model, precision, flags) &&
                            max_steps >= 0 && (uint32_t)max_steps <= ddm::WIRE_MAX_STEPS;
    // How a host-destined batch travels: one launch + one copy, plain chunks (kernel of chunk i+1 over the copy of chunk i),
    // or compact records + host decode.  A caller's ddm_set_pipeline(min_rows) overrides the measured defaults.
    const int64_t rows_total = n_datasets * n_trials;
    bool stream = false, want_compact = compact_ok;
    if (out_host && n_trials > 0 && n_datasets >= 2 && !(flags & DDM_FLAG_KEEP_STEPS) && !ctx->dbg_on) {
        if (ctx->tune_pipeline_min_rows >= 0) {
            stream = rows_total >= ctx->tune_pipeline_min_rows;
        } else {
            const int k = kind_of(model);
            const bool basic_cols = (k == ddm::KIND_FIXED || k == ddm::KIND_DRIFT);
            if (compact_ok && rows_total >= (basic_cols ? kCompactMinRows : kCompactMinRowsExt)) {
                stream = true;
            } else if (rows_total >= kPipelineMinRows) {
                stream = true;
                want_compact = false;
            } else if (rows_total >= kPinnedPipelineMinRows) {
                cudaPointerAttributes attr{};
                const bool pinned = cudaPointerGetAttributes(&attr, out_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
                cudaGetLastError();  // an unregistered pointer is not an error worth keeping
                stream = pinned;
                want_compact = false;
            }
        }
    }
    if (stream) {
        DeviceGuard g(ctx->device);
        return run_pipelined(ctx, params, n_trials, dt, max_steps, seed, dataset_offset, precision, flags, out_host, want_compact);
    }
    if (n_datasets > 0) {
        DeviceGuard g(ctx->device);
        DDM_CUDA(ctx, cudaMemcpyAsync(ctx->params.p, params, (size_t)n_datasets * n_params * sizeof(double), cudaMemcpyHostToDevice,
                                      ctx->stream));
    }
    rc = ddm_run(ctx, n_trials, dt, max_steps, seed, dataset_offset, precision, flags);
    if (rc) return rc;
    if (out_host) return ddm_download(ctx, out_host);
    return DDM_OK;
}

DDM_API int ddm_simulate_trialwise(ddm_ctx *ctx, const int32_t *group, const double *bound, const double *group_params,
                                   int64_t n, int n_groups, double dt, int max_steps, uint64_t seed,
                                   uint64_t trial_offset, int precision, int flags, void *out_host) {
    if (!ctx) return DDM_ERR_INVALID;
    if (n < 0 || n_groups < 0) return fail(ctx, DDM_ERR_INVALID, "negative shape");
    if (n > 0 && (!group || !bound || !group_params)) return fail(ctx, DDM_ERR_INVALID, "NULL input");
    {
        // imputation_from_stahl_not_scaled.py:124-125 raises ValueError; NaN passes there too (NaN < 0 is False).
        // One pass over the inputs, shared among the host threads for large batches; the first offender is reported.
        struct Scan {
            const int32_t *group;
            const double *bound;
            int64_t n;
            int n_groups;
            std::atomic<int64_t> bad_bound, bad_group;
        } scan{group, bound, n, n_groups, {INT64_MAX}, {INT64_MAX}};
        auto slice = [](const void *arg, int id, int nthr) {
            Scan &sc = *const_cast<Scan *>(static_cast<const Scan *>(arg));
            const int64_t per = (sc.n + nthr - 1) / nthr, lo = per * id, hi = std::min<int64_t>(sc.n, lo + per);
            for (int64_t i = lo; i < hi; i++) {
                if (sc.bound[i] < 0) {
                    int64_t cur = sc.bad_bound.load(std::memory_order_relaxed);
                    while (i < cur && !sc.bad_bound.compare_exchange_weak(cur, i)) {}
                    break;
                }
                if (sc.group[i] < 0 || sc.group[i] >= sc.n_groups) {
                    int64_t cur = sc.bad_group.load(std::memory_order_relaxed);
                    while (i < cur && !sc.bad_group.compare_exchange_weak(cur, i)) {}
                    break;
                }
            }
        };
        if (n >= (1 << 20) && ctx->tune_host_decode >= 0) {
            int rc = ensure_workers(ctx);
            if (rc) return rc;
            ddm::host_workers_run(ctx->workers, slice, &scan);
        } else {
            slice(&scan, 0, 1);
        }
        const int64_t bb = scan.bad_bound.load(), bg = scan.bad_group.load();
        if (bb != INT64_MAX && bb < bg)
            return fail(ctx, DDM_ERR_NEGATIVE_BOUND, "Trial-level boundary cannot be less than zero (trial %lld: %g)",
                        (long long)bb, bound[bb]);
        if (bg != INT64_MAX)
            return fail(ctx, DDM_ERR_INVALID, "group[%lld] = %d outside [0,%d)", (long long)bg, group[bg], n_groups);
        if (bb != INT64_MAX)
            return fail(ctx, DDM_ERR_NEGATIVE_BOUND, "Trial-level boundary cannot be less than zero (trial %lld: %g)",
                        (long long)bb, bound[bb]);
    }
    ctx->trialwise_degenerate = false;
    for (int gidx = 0; gidx < n_groups; gidx++) {
        const double dc = group_params[(size_t)gidx * 4 + 3];
        if (!(dc > 1e-30) || !std::isfinite(dc)) ctx->trialwise_degenerate = true;  // no noise unit: reference formulas
    }
    DeviceGuard g(ctx->device);
    DDM_CUDA(ctx, ctx->group.reserve(n ? (size_t)n : 1));
    DDM_CUDA(ctx, ctx->bound.reserve(n ? (size_t)n : 1));
    DDM_CUDA(ctx, ctx->params.reserve(n_groups ? (size_t)n_groups * 4 : 1));
    ctx->have_params = false;  // the params arena now holds group parameters
    if (n) {
        DDM_CUDA(ctx, cudaMemcpyAsync(ctx->group.p, group, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        DDM_CUDA(ctx, cudaMemcpyAsync(ctx->bound.p, bound, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        DDM_CUDA(ctx, cudaMemcpyAsync(ctx->params.p, group_params, (size_t)n_groups * 4 * sizeof(double), cudaMemcpyHostToDevice,
                                      ctx->stream));
    }
    int rc = run_common(ctx, DDM_MODEL_TRIALWISE, 1, n, dt, max_steps, seed, 0, trial_offset, precision, flags, n_groups);
    if (rc) return rc;
    if (out_host) return ddm_download(ctx, out_host);
    return DDM_OK;
}

// Exact rejection sampler: pyhddmjagsutils.py:47-176 (simulratcliff).
DDM_API int ddm_simulate_exact(ddm_ctx *ctx, const double *params, int64_t n_datasets, int64_t n_trials, uint64_t seed,
                               uint64_t dataset_offset, double *out_host) {
    if (!ctx) return DDM_ERR_INVALID;
    if (n_trials < 0 || n_datasets < 0) return fail(ctx, DDM_ERR_INVALID, "negative shape");
    if (n_trials > 0xffffffffLL || n_datasets > 0xffffffffLL)
        return fail(ctx, DDM_ERR_INVALID, "shape exceeds the 32-bit Philox counter words");
    uint32_t ds_lo = 0, ds_hi = 0;
    if (int rc0 = split_index(ctx, dataset_offset, n_datasets, &ds_lo, &ds_hi)) return rc0;
    if (n_datasets > 0 && !params) return fail(ctx, DDM_ERR_INVALID, "params is NULL");
    for (int64_t d = 0; d < n_datasets; d++) {
        const double *p = params + (size_t)d * 8;
        for (int j = 0; j < 8; j++)
            if (!std::isfinite(p[j])) return fail(ctx, DDM_ERR_INVALID, "dataset %lld: parameter %d is not finite", (long long)d, j);
        if (!(p[0] > 0.0) || !(p[7] > 0.0))
            return fail(ctx, DDM_ERR_INVALID, "dataset %lld: Alpha and Varsigma must be positive (got %g, %g)", (long long)d, p[0], p[7]);
        const double b_lo = p[3] - std::fabs(p[5]) / 2, b_hi = p[3] + std::fabs(p[5]) / 2;
        if (b_lo < 0.0 || b_hi > 1.0)
            return fail(ctx, DDM_ERR_INVALID, "dataset %lld: start point Beta +- rangeBeta/2 leaves [0, 1]", (long long)d);
    }
    DeviceGuard g(ctx->device);
    const size_t np = (size_t)n_datasets * 8;
    DDM_CUDA(ctx, ctx->params.reserve(np ? np : 1));
    if (np) DDM_CUDA(ctx, cudaMemcpyAsync(ctx->params.p, params, np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->have_params = false;  // the params arena now holds the sampler's parameters
    const int64_t rows = n_datasets * n_trials;
    const size_t out_bytes = (size_t)rows * sizeof(double);
    int rc = ensure_output(ctx, out_bytes);
    if (rc) return rc;
    ddm_stats st{};
    st.n_trials = (uint64_t)rows;
    DDM_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT), ctx->stream));
    DDM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    const ddm::PhiloxKey key = ddm::make_philox_key((uint32_t)seed, (uint32_t)(seed >> 32), ds_hi);
    DDM_CUDA(ctx, ddm::launch_exact_sampler(ctx->params.p, static_cast<double *>(ctx->out), ctx->counters + 1, (uint32_t)n_datasets,
                                            (uint32_t)n_trials, ds_lo, 0u, key, ctx->stream));
    if (rows) st.kernel_launches++;
    st.grid = (int)(((uint64_t)rows + 127) / 128);
    st.block = 128;
    DDM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, ctx->counters, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats = st;
    ctx->stats_pending = true;
    ctx->have_run = true;
    ctx->out64 = true;
    ctx->out_bytes = out_bytes;
    ctx->have_steps = false;
    ctx->run_rows = rows;
    ctx->run_datasets = n_datasets;
    ctx->run_trials = n_trials;
    ctx->run_trialwise = false;
    ctx->run_cols = 1;
    ctx->run_model = kRunModelExact;
    ctx->out_resident = true;
    if (out_host) return ddm_download(ctx, out_host);
    return DDM_OK;
}

// Evidence-path variants: retired_models/basic_ddm_dc_evidence.py:87-151, basic_ddm_dc_evidence2.py:83-150,
// basic_ddm_dc_evidence_no_noise2.py:82-147.
DDM_API int ddm_simulate_evidence(ddm_ctx *ctx, const double *params, int64_t n_datasets, int64_t n_trials, int n_obs,
                                  int standardize, double dt, int max_steps, uint64_t seed, uint64_t dataset_offset,
                                  int precision, int flags, void *out_host) {
    if (!ctx) return DDM_ERR_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, DDM_ERR_INVALID, "precision must be 32 or 64, got %d", precision);
    if (!(dt > 0.0) || !std::isfinite(dt)) return fail(ctx, DDM_ERR_INVALID, "dt must be positive and finite");
    if (max_steps < 0 || n_trials < 0 || n_datasets < 0) return fail(ctx, DDM_ERR_INVALID, "negative argument");
    if (n_obs < 1 || n_obs > 32 * 6 * 4) return fail(ctx, DDM_ERR_INVALID, "n_obs must be in [1, 768], got %d", n_obs);
    if (standardize < 0 || standardize > 2) return fail(ctx, DDM_ERR_INVALID, "standardize must be 0, 1 or 2");
    if (n_trials > 0xffffffffLL || n_datasets > 0xffffffffLL)
        return fail(ctx, DDM_ERR_INVALID, "shape exceeds the 32-bit Philox counter words");
    uint32_t ds_lo = 0, ds_hi = 0;
    if (int rc0 = split_index(ctx, dataset_offset, n_datasets, &ds_lo, &ds_hi)) return rc0;
    if (n_datasets > 0 && !params) return fail(ctx, DDM_ERR_INVALID, "params is NULL");
    const int64_t rows = n_datasets * n_trials;
    if (ctx->dbg_on && precision != 64) return fail(ctx, DDM_ERR_INVALID, "shared-increment mode needs precision 64 for the evidence model");
    if (ctx->dbg_on && ctx->dbg_trials != rows)
        return fail(ctx, DDM_ERR_INVALID, "shared-increment buffer was set for %lld trials, run has %lld",
                    (long long)ctx->dbg_trials, (long long)rows);
    // the recording kernel holds a trial's final state, which equals the state at max_steps only when the
    // observation window ends before max_steps (always so in the reference: .2 s or .4 s of 4 s)
    if (precision == 32 && max_steps < n_obs) precision = 64;
    for (int64_t d = 0; d < n_datasets && precision == 32; d++) {
        const double dc = params[(size_t)d * 6 + 4];
        if (!(dc > 1e-30) || !std::isfinite(dc)) precision = 64;  // no noise unit: the validation kernel keeps the reference's formulas
    }
    DeviceGuard g(ctx->device);
    const size_t np = (size_t)n_datasets * 6;
    DDM_CUDA(ctx, ctx->params.reserve(np ? np : 1));
    if (np) DDM_CUDA(ctx, cudaMemcpyAsync(ctx->params.p, params, np * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->have_params = false;  // the params arena now holds evidence parameters
    const bool out64 = !(flags & DDM_FLAG_OUT_F32);
    const uint32_t cols = 2u + (uint32_t)n_obs;
    const size_t total = (size_t)rows * cols;
    const size_t out_bytes = total * (out64 ? 8 : 4);
    int rc = ensure_output(ctx, out_bytes);
    if (rc) return rc;

    ddm::EvidenceArgs a{};
    a.params = ctx->params.p;
    a.dbg_z = ctx->dbg_on ? ctx->dbg_z.p : nullptr;
    a.dbg_off = ctx->dbg_on ? ctx->dbg_off.p : nullptr;
    a.dbg_n = ctx->dbg_on ? ctx->dbg_n : 0;
    a.out = ctx->out;
    a.work_counter = ctx->counters;
    a.stats = ctx->counters + 1;
    a.n_datasets = (uint32_t)n_datasets;
    a.n_trials = (uint32_t)n_trials;
    a.n_obs = (uint32_t)n_obs;
    a.tiles_per_dataset = (uint32_t)((n_trials + 31) / 32);
    a.n_items = (uint64_t)a.tiles_per_dataset * (uint64_t)n_datasets;
    if (a.n_items > 0xffffffffULL) return fail(ctx, DDM_ERR_INVALID, "too many trial tiles for one launch");
    a.dataset_offset = ds_lo;
    a.trial_offset = 0;
    a.max_steps = (uint32_t)max_steps;
    a.mode = standardize;
    a.flags = flags;
    a.key = ddm::make_philox_key((uint32_t)seed, (uint32_t)(seed >> 32), ds_hi);
    a.dt = dt;
    a.sqrt_dt = std::sqrt(dt);
    if (standardize == 2) {
        DDM_CUDA(ctx, ctx->ev_means.reserve(rows ? (size_t)rows : 1));
        DDM_CUDA(ctx, ctx->ev_ds_stats.reserve(n_datasets ? (size_t)n_datasets * 2 : 2));
        a.path_means = ctx->ev_means.p;
    }
    ddm_stats st{};
    st.n_trials = (uint64_t)rows;
    DDM_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT), ctx->stream));
    if (rows > 0) {
        DDM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        if (precision == 32) {
            // (1) step + record with the persistent refill kernel (the basic model's lanes and counters)
            DDM_CUDA(ctx, ctx->dconst.reserve((size_t)n_datasets));
            DDM_CUDA(ctx, ctx->ev_pairs.reserve((size_t)rows));  // 8 bytes per trial: (steps, choice), final state
            a.rec_g = ddm::evidence_lanes_per_trial(a.n_obs);
            a.rec_stride = ddm::evidence_rec_stride(a.n_obs);
            DDM_CUDA(ctx, ctx->ev_path.reserve((size_t)rows * a.rec_stride));
            DDM_CUDA(ctx, ddm::launch_prep(ctx->params.p, ctx->dconst.p, (uint32_t)n_datasets, 6u, DDM_MODEL_BASIC, dt, ctx->stream));
            ddm::RunArgs r{};
            r.dconst = ctx->dconst.p;
            r.params = ctx->params.p;
            r.out = ctx->ev_pairs.p;
            r.rec_path = ctx->ev_path.p;
            r.rec_stride = a.rec_stride;
            r.rec_g = a.rec_g;
            r.n_obs = a.n_obs;
            r.work_counter = ctx->counters;
            r.stats = ctx->counters + 1;
            r.n_datasets = a.n_datasets;
            r.n_trials = a.n_trials;
            r.n_params = 6;
            r.dataset_offset = a.dataset_offset;
            r.trial_offset = 0;
            r.key = a.key;
            r.max_steps = a.max_steps;
            r.model = DDM_MODEL_BASIC;
            r.flags = 0;
            r.dt = dt;
            r.sqrt_dt = a.sqrt_dt;
            uint32_t tile = ctx->tune_tile > 0 ? (uint32_t)ctx->tune_tile : 64u;
            if (tile > a.n_trials) tile = a.n_trials;
            r.tile = tile;
            r.tiles_per_dataset = (a.n_trials + tile - 1) / tile;
            r.n_items = (uint64_t)r.tiles_per_dataset * a.n_datasets;
            if (r.n_items > 0xffffffffULL) return fail(ctx, DDM_ERR_INVALID, "too many trial tiles for one launch");
            r.refill_threshold = ctx->tune_threshold > 0 ? ctx->tune_threshold : default_refill_threshold(dt);
            const int block = ddm::record_block_size();
            int per_sm = ddm::record_max_blocks_per_sm(block);
            if (per_sm <= 0) return fail(ctx, DDM_ERR_CUDA, "occupancy query failed for the recording kernel");
            uint64_t grid = (uint64_t)ctx->sm_count * per_sm;
            const uint64_t need = ((uint64_t)rows + block - 1) / block;
            if (grid > need) grid = need;
            if (grid < 1) grid = 1;
            DDM_CUDA(ctx, ddm::launch_record(r, (int)grid, block, ctx->stream));
            // (2) eight lanes per trial: noise, standardisation, row stores
            a.rec_path = ctx->ev_path.p;
            a.rec_meta = reinterpret_cast<const uint2 *>(ctx->ev_pairs.p);
            a.dconst = ctx->dconst.p;
            DDM_CUDA(ctx, ddm::launch_evidence_post(a, out64, (uint64_t)rows, ctx->sm_count, ctx->stream));
            st.kernel_launches += 3;
            st.grid = (int)grid;
            st.block = block;
            st.used_persistent = 1;
            st.scheduler = 1;
            st.refill_threshold = r.refill_threshold;
            st.tile = (int)tile;
            if (standardize == 2) {
                DDM_CUDA(ctx, ddm::launch_evidence_dataset_stats(ctx->ev_means.p, ctx->ev_ds_stats.p, a.n_datasets, a.n_trials, ctx->stream));
                DDM_CUDA(ctx, ddm::launch_evidence_finalize(ctx->out, out64, ctx->out, out64, ctx->ev_ds_stats.p, total, cols,
                                                            a.n_trials, true, ctx->stream));
                st.kernel_launches += 2;
            }
        } else {
            DDM_CUDA(ctx, ctx->ev_scratch.reserve(total));
            a.scratch = ctx->ev_scratch.p;
            DDM_CUDA(ctx, ddm::launch_evidence_generic(a, ctx->dbg_on, (uint64_t)rows, ctx->stream));
            st.kernel_launches++;
            st.grid = (int)((rows + 127) / 128);
            st.block = 128;
            if (standardize == 2) {
                DDM_CUDA(ctx, ddm::launch_evidence_dataset_stats(ctx->ev_means.p, ctx->ev_ds_stats.p, a.n_datasets, a.n_trials, ctx->stream));
                st.kernel_launches++;
            }
            DDM_CUDA(ctx, ddm::launch_evidence_finalize(ctx->ev_scratch.p, true, ctx->out, out64, ctx->ev_ds_stats.p, total, cols,
                                                        a.n_trials, standardize == 2, ctx->stream));
            st.kernel_launches++;
        }
        DDM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    }
    DDM_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, ctx->counters, sizeof(unsigned long long) * (1 + ddm::STAT_COUNT),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats = st;
    ctx->stats_pending = true;
    ctx->have_run = true;
    ctx->out64 = out64;
    ctx->out_bytes = out_bytes;
    ctx->have_steps = false;
    ctx->run_rows = rows;
    ctx->run_datasets = n_datasets;
    ctx->run_trials = n_trials;
    ctx->run_trialwise = false;
    ctx->run_cols = cols;
    ctx->run_model = kRunModelEvidence;
    ctx->out_resident = true;
    if (out_host) return ddm_download(ctx, out_host);
    return DDM_OK;
}

DDM_API int ddm_last_steps(ddm_ctx *ctx, int32_t *steps_host) {
    if (!ctx) return DDM_ERR_INVALID;
    if (!ctx->have_run || !ctx->have_steps) return fail(ctx, DDM_ERR_STATE, "last run did not keep step counts (DDM_FLAG_KEEP_STEPS)");
    DeviceGuard g(ctx->device);
    if (ctx->run_rows)
        DDM_CUDA(ctx, cudaMemcpyAsync(steps_host, ctx->steps.p, (size_t)ctx->run_rows * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DDM_OK;
}

DDM_API int ddm_last_stats(ddm_ctx *ctx, ddm_stats *out) {
    if (!ctx || !out) return DDM_ERR_INVALID;
    if (!ctx->have_run) return fail(ctx, DDM_ERR_STATE, "no run yet");
    DeviceGuard g(ctx->device);
    int rc = finish_stats(ctx);
    if (rc) return rc;
    *out = ctx->stats;
    return DDM_OK;
}

DDM_API int ddm_last_output_histogram(ddm_ctx *ctx, int n_bins, double rt_max, uint64_t *hist_host) {
    if (!ctx || !hist_host) return DDM_ERR_INVALID;
    if (n_bins < 1 || n_bins > 8192 || !(rt_max > 0.0)) return fail(ctx, DDM_ERR_INVALID, "need 1 <= n_bins <= 8192 and rt_max > 0");
    if (!ctx->have_run || !ctx->out || !ctx->out_resident) return fail(ctx, DDM_ERR_STATE, "no output resident");
    if (ctx->run_model == DDM_MODEL_GENERAL) return fail(ctx, DDM_ERR_STATE, "histogram is not defined for DDM_MODEL_GENERAL rows");
    DeviceGuard g(ctx->device);
    const size_t n_cells = 2 * (size_t)n_bins + 2;
    DDM_CUDA(ctx, ctx->hist.reserve(n_cells));
    DDM_CUDA(ctx, cudaMemsetAsync(ctx->hist.p, 0, n_cells * sizeof(unsigned long long), ctx->stream));
    const bool basic = ctx->run_model == DDM_MODEL_BASIC || ctx->run_model == DDM_MODEL_ETA || ctx->run_model == kRunModelEvidence;
    DDM_CUDA(ctx, ddm::launch_rt_histogram(ctx->out, ctx->out64, (uint64_t)ctx->run_rows, (uint32_t)ctx->run_cols, basic, (uint32_t)n_bins,
                                           rt_max, ctx->hist.p, ctx->sm_count, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(hist_host, ctx->hist.p, n_cells * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DDM_OK;
}

// C5 as SURVEY.md section 8d specifies it: host parameters in, the batch simulated and reduced on the device, only the
// histogram comes back.  One call = H2D of the parameters, prep + simulator kernel (float32 rows, resident),
// histogram kernel, D2H of 2 n_bins + 2 counters.
DDM_API int ddm_simulate_histogram(ddm_ctx *ctx, int model, const double *params, int64_t n_datasets, int n_params,
                                   int64_t n_trials, double dt, int max_steps, uint64_t seed, uint64_t dataset_offset,
                                   int precision, int flags, int n_bins, double rt_max, uint64_t *hist_host) {
    if (!ctx || !hist_host) return DDM_ERR_INVALID;
    if (n_bins < 1 || n_bins > 8192 || !(rt_max > 0.0)) return fail(ctx, DDM_ERR_INVALID, "need 1 <= n_bins <= 8192 and rt_max > 0");
    if (model == DDM_MODEL_GENERAL || model == DDM_MODEL_TRIALWISE)
        return fail(ctx, DDM_ERR_INVALID, "ddm_simulate_histogram takes the two-column dataset-wise models");
    const int run_flags = (flags | DDM_FLAG_OUT_F32) & ~DDM_FLAG_KEEP_STEPS;
    // ddm_set_pipeline governs this path as it does ddm_simulate's: min_rows >= 0 replaces the threshold (a huge one
    // switches the chunking off, 0 forces it), chunk_rows > 0 the smallest chunk
    const int64_t stream_from = ctx->tune_pipeline_min_rows >= 0 ? ctx->tune_pipeline_min_rows : kHistStreamMinRows;
    if (n_datasets > 0 && n_trials > 0 && n_datasets * n_trials >= stream_from && !ctx->dbg_on) {
        int rc = upload_params_impl(ctx, model, params, n_datasets, n_params, false);
        if (rc) return rc;
        if (takes_persistent_kernel(ctx, model, precision, run_flags) && max_steps >= 0 && (uint32_t)max_steps <= ddm::TILE_MAX_STEPS) {
            DeviceGuard g(ctx->device);
            return run_histogram_streamed(ctx, params, n_trials, dt, max_steps, seed, dataset_offset, precision, run_flags, n_bins, rt_max,
                                          hist_host);
        }
    }
    int rc = ddm_upload_params(ctx, model, params, n_datasets, n_params);
    if (rc) return rc;
    rc = ddm_run(ctx, n_trials, dt, max_steps, seed, dataset_offset, precision, run_flags);
    if (rc) return rc;
    if (n_datasets * n_trials == 0) {
        std::memset(hist_host, 0, (2 * (size_t)n_bins + 2) * sizeof(uint64_t));
        return DDM_OK;
    }
    return ddm_last_output_histogram(ctx, n_bins, rt_max, hist_host);
}

DDM_API int ddm_last_output_device_ptr(ddm_ctx *ctx, void **ptr, size_t *bytes) {
    if (!ctx || !ptr) return DDM_ERR_INVALID;
    if (!ctx->have_run || !ctx->out || !ctx->out_resident) return fail(ctx, DDM_ERR_STATE, "no output resident");
    *ptr = ctx->out;
    if (bytes) *bytes = ctx->out_bytes;
    return DDM_OK;
}

DDM_API int ddm_last_output_dlpack(ddm_ctx *ctx, struct DLManagedTensor **out) {
    if (!ctx || !out) return DDM_ERR_INVALID;
    if (!ctx->have_run || !ctx->out || !ctx->out_resident)
        return fail(ctx, DDM_ERR_STATE, "no output resident (already handed off, or the last run streamed to the host)");
    DeviceGuard g(ctx->device);
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    auto *h = new (std::nothrow) DlpackHolder();
    if (!h) return fail(ctx, DDM_ERR_NOMEM, "out of host memory");
    h->buf = ctx->out;
    h->cap = ctx->out_cap;
    h->pool = ctx->pool;
    ctx->pool->refs.fetch_add(1);
    DLTensor &t = h->t.dl_tensor;
    t.data = ctx->out;
    t.device.device_type = kDLCUDA;
    t.device.device_id = ctx->device;
    t.dtype.code = kDLFloat;
    t.dtype.bits = ctx->out64 ? 64 : 32;
    t.dtype.lanes = 1;
    if (ctx->run_trialwise) {
        t.ndim = 2;
        h->shape[0] = ctx->run_rows;
        h->shape[1] = 2;
    } else {
        t.ndim = 3;
        h->shape[0] = ctx->run_datasets;
        h->shape[1] = ctx->run_trials;
        h->shape[2] = ctx->run_cols;
    }
    t.shape = h->shape;
    t.strides = nullptr;
    t.byte_offset = 0;
    h->t.manager_ctx = h;
    h->t.deleter = dlpack_deleter;
    ctx->out = nullptr;  // ownership moved to the consumer
    ctx->out_cap = 0;
    *out = &h->t;
    return DDM_OK;
}

// One online-training batch in one call: what bf.simulation.GenerativeModel(prior, simulator)(batch_size) produces
// (basic_ddm_dc.py:129-133, single_trial_alpha_not_scaled.py:270-274) -- prior draws on the device, simulated where they
// are, the draws copied out for the trainer's targets, the batch handed over as DLPack -- with everything enqueued
// before the one stream synchronisation (the three-call sequence ddm_draw_prior / ddm_run / ddm_last_output_dlpack
// waits for the GPU twice and crosses the FFI three times; at 64 x 500 trials that is a third of the batch's latency).
DDM_API int ddm_training_batch(ddm_ctx *ctx, int prior, int64_t n_draws, int64_t n_trials, double dt, int max_steps, uint64_t seed,
                               uint64_t draw_offset, int flags, double *params_host, struct DLManagedTensor **out) {
    if (!ctx || !out) return DDM_ERR_INVALID;
    if (prior == DDM_PRIOR_EVIDENCE) return fail(ctx, DDM_ERR_INVALID, "ddm_training_batch takes the dataset-wise models' priors");
    if (n_draws <= 0 || n_trials <= 0) return fail(ctx, DDM_ERR_INVALID, "ddm_training_batch needs n_draws >= 1 and n_trials >= 1");
    int rc = ddm_draw_prior(ctx, prior, n_draws, seed, draw_offset, nullptr);
    if (rc) return rc;
    rc = ddm_run(ctx, n_trials, dt, max_steps, seed, draw_offset, 32, flags);
    if (rc) return rc;
    const size_t n = (size_t)n_draws * (size_t)ctx->n_params;
    if (params_host) {
        DeviceGuard g(ctx->device);
        if (ctx->train_stage_cap < n) {
            if (ctx->train_stage) cudaFreeHost(ctx->train_stage);
            ctx->train_stage = nullptr;
            ctx->train_stage_cap = 0;
            DDM_CUDA(ctx, cudaHostAlloc(&ctx->train_stage, n * sizeof(double), cudaHostAllocDefault));
            ctx->train_stage_cap = n;
        }
        DDM_CUDA(ctx, cudaMemcpyAsync(ctx->train_stage, ctx->params.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    rc = ddm_last_output_dlpack(ctx, out);  // synchronises the stream
    if (rc) return rc;
    if (params_host) std::memcpy(params_host, ctx->train_stage, n * sizeof(double));
    return DDM_OK;
}

// ---- parity hooks -------------------------------------------------------------------------
DDM_API int ddm_set_normals_debug(ddm_ctx *ctx, const double *z, size_t n, const int64_t *offsets, int64_t n_trials) {
    if (!ctx) return DDM_ERR_INVALID;
    if (!z) {
        ctx->dbg_on = false;
        return DDM_OK;
    }
    if (!offsets || n_trials < 0) return fail(ctx, DDM_ERR_INVALID, "offsets required");
    for (int64_t i = 0; i < n_trials; i++)
        if (offsets[i] < 0 || (size_t)offsets[i] > n) return fail(ctx, DDM_ERR_INVALID, "offsets[%lld] outside the buffer", (long long)i);
    DeviceGuard g(ctx->device);
    DDM_CUDA(ctx, ctx->dbg_z.reserve(n ? n : 1));
    DDM_CUDA(ctx, ctx->dbg_off.reserve(n_trials ? (size_t)n_trials : 1));
    if (n) DDM_CUDA(ctx, cudaMemcpyAsync(ctx->dbg_z.p, z, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (n_trials)
        DDM_CUDA(ctx, cudaMemcpyAsync(ctx->dbg_off.p, offsets, (size_t)n_trials * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->dbg_on = true;
    ctx->dbg_n = n;
    ctx->dbg_trials = n_trials;
    return DDM_OK;
}

DDM_API int ddm_export_normals(ddm_ctx *ctx, uint64_t seed, uint64_t dataset, uint32_t trial, uint32_t stream,
                               uint32_t first, uint32_t count, int precision, double *out_host) {
    if (!ctx) return DDM_ERR_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, DDM_ERR_INVALID, "precision must be 32 or 64");
    if (stream > 1) return fail(ctx, DDM_ERR_INVALID, "stream must be 0 (step) or 1 (aux)");
    if (count == 0) return DDM_OK;
    if (!out_host) return fail(ctx, DDM_ERR_INVALID, "out_host is NULL");
    DeviceGuard g(ctx->device);
    DDM_CUDA(ctx, ctx->export_buf.reserve(count));
    if ((dataset >> 32) >= (1ULL << 24)) return fail(ctx, DDM_ERR_INVALID, "global dataset index must stay below 2^56");
    const ddm::PhiloxKey key = ddm::make_philox_key((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(dataset >> 32));
    DDM_CUDA(ctx, ddm::launch_export_normals(key, (uint32_t)dataset, trial, stream, first, count, precision == 64, ctx->export_buf.p,
                                             ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->export_buf.p, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DDM_OK;
}

DDM_API int ddm_normals_histogram(ddm_ctx *ctx, uint64_t seed, uint64_t n_normals, int n_bins_abs, double z_max, int n_bins_angle,
                                  uint64_t *hist_host, double *moments_host) {
    if (!ctx || !hist_host || !moments_host) return DDM_ERR_INVALID;
    if (n_bins_abs < 1 || n_bins_abs > 4096 || n_bins_angle < 1 || n_bins_angle > 4096 || !(z_max > 0.0))
        return fail(ctx, DDM_ERR_INVALID, "need 1 <= bins <= 4096 and z_max > 0");
    DeviceGuard g(ctx->device);
    const size_t n_cells = (size_t)n_bins_abs + 1 + (size_t)n_bins_angle;
    DDM_CUDA(ctx, ctx->hist.reserve(n_cells + 4));  // histogram + four double moments behind it
    DDM_CUDA(ctx, cudaMemsetAsync(ctx->hist.p, 0, (n_cells + 4) * sizeof(unsigned long long), ctx->stream));
    double *mom = reinterpret_cast<double *>(ctx->hist.p + n_cells);
    const ddm::PhiloxKey key = ddm::make_philox_key((uint32_t)seed, (uint32_t)(seed >> 32));
    DDM_CUDA(ctx, ddm::launch_normals_histogram(key, (n_normals + 5) / 6, (uint32_t)n_bins_abs, z_max, (uint32_t)n_bins_angle, ctx->hist.p, mom,
                                                ctx->sm_count, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(hist_host, ctx->hist.p, n_cells * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(moments_host, mom, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DDM_OK;
}

DDM_API int64_t ddm_pipeline_chunks(int64_t n_datasets, int64_t n_trials, int64_t chunk_rows, int64_t *first, int64_t *count,
                                    int64_t capacity) {
    if (n_datasets < 0 || n_trials < 0) return DDM_ERR_INVALID;
    const auto chunks = pipeline_chunks(n_datasets, n_trials, chunk_rows);
    for (size_t i = 0; i < chunks.size() && (int64_t)i < capacity; i++) {
        if (first) first[i] = chunks[i].first;
        if (count) count[i] = chunks[i].second;
    }
    return (int64_t)chunks.size();
}

DDM_API int64_t ddm_histogram_chunks(int64_t n_datasets, int64_t n_trials, int64_t min_chunk_rows, int64_t *first, int64_t *count,
                                     int64_t capacity) {
    if (n_datasets < 0 || n_trials < 0) return -1;
    const auto chunks = histogram_chunks(n_datasets, n_trials, min_chunk_rows);
    for (int64_t i = 0; i < (int64_t)chunks.size() && i < capacity; i++) {
        if (first) first[i] = chunks[i].first;
        if (count) count[i] = chunks[i].second;
    }
    return (int64_t)chunks.size();
}

DDM_API int ddm_wire_decode_host(const void *wire, void *out_host, const double *params, int n_params, int64_t n_datasets,
                                 int64_t n_trials, double dt, int basic_columns, int flags, int n_threads) {
    if (n_datasets < 0 || n_trials < 0 || n_params < 4 || n_threads < 1 || n_threads > 256) return DDM_ERR_INVALID;
    if (n_datasets * n_trials > 0 && (!wire || !out_host || !params)) return DDM_ERR_INVALID;
    ddm::HostWorkers *w = ddm::host_workers_create(n_threads);
    ddm::WireDecode job;
    job.wire = wire;
    job.out = out_host;
    job.params = params;
    job.n_params = n_params;
    job.tau_col = 3;
    job.n_datasets = n_datasets;
    job.n_trials = n_trials;
    job.dt = dt;
    job.basic = basic_columns != 0;
    job.out64 = !(flags & DDM_FLAG_OUT_F32);
    job.timeout_choice_one = (flags & DDM_FLAG_TIMEOUT_CHOICE_ONE) != 0;
    ddm::wire_decode(w, job);
    ddm::host_workers_destroy(w);
    return DDM_OK;
}

DDM_API int ddm_host_stream_peak(int n_threads, size_t bytes, double *bytes_per_s) {
    if (!bytes_per_s || n_threads < 0 || n_threads > 256 || bytes < (1u << 20)) return DDM_ERR_INVALID;
    if (n_threads == 0) {
        int gpus = 1;
        if (cudaGetDeviceCount(&gpus) != cudaSuccess) gpus = 1;
        n_threads = ddm::host_workers_default_count(gpus);
    }
    void *buf = nullptr;
    if (posix_memalign(&buf, 4096, bytes) != 0 || !buf) return DDM_ERR_NOMEM;
    ddm::HostWorkers *w = ddm::host_workers_create(n_threads);
    ddm::host_workers_begin(w);
    *bytes_per_s = ddm::host_stream_store_rate(w, buf, bytes, 5);
    ddm::host_workers_end(w);
    ddm::host_workers_destroy(w);
    free(buf);
    return DDM_OK;
}

DDM_API int ddm_philox4x32(ddm_ctx *ctx, const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4, int64_t n_blocks) {
    if (!ctx) return DDM_ERR_INVALID;
    if (n_blocks < 0) return fail(ctx, DDM_ERR_INVALID, "n_blocks < 0");
    if (n_blocks == 0) return DDM_OK;
    if (!ctr4 || !key2 || !out4) return fail(ctx, DDM_ERR_INVALID, "NULL argument");
    DeviceGuard g(ctx->device);
    const size_t n = (size_t)n_blocks;
    DDM_CUDA(ctx, ctx->philox_buf.reserve(n * 10));
    uint32_t *d_ctr = ctx->philox_buf.p, *d_key = d_ctr + 4 * n, *d_out = d_key + 2 * n;
    DDM_CUDA(ctx, cudaMemcpyAsync(d_ctr, ctr4, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(d_key, key2, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    DDM_CUDA(ctx, ddm::launch_philox_blocks(d_ctr, d_key, d_out, n_blocks, ctx->stream));
    DDM_CUDA(ctx, cudaMemcpyAsync(out4, d_out, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    DDM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DDM_OK;
}

// ---- measurement ------------------------------------------------------------------------------
DDM_API int ddm_microbench(ddm_ctx *ctx, int which, int iters, double *inst_per_s, double *sm_hz) {
    if (!ctx) return DDM_ERR_INVALID;
    if (which < 0 || which >= DDM_MB_COUNT) return fail(ctx, DDM_ERR_INVALID, "unknown micro-benchmark %d", which);
    if (iters <= 0) iters = 4096;
    DeviceGuard g(ctx->device);
    double ips = 0, hz = 0;
    cudaError_t e = ddm::run_microbench(which, iters, ctx->sm_count, ctx->stream, &ips, &hz);
    if (e != cudaSuccess) return fail(ctx, DDM_ERR_CUDA, "microbench %d: %s", which, cudaGetErrorString(e));
    if (inst_per_s) *inst_per_s = ips;
    if (sm_hz) *sm_hz = hz;
    return DDM_OK;
}

DDM_API int ddm_host_alloc(size_t bytes, void **ptr) {
    if (!ptr) return DDM_ERR_INVALID;
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaMallocHost(") + std::to_string(bytes) + "): " + cudaGetErrorName(e) + ": " + cudaGetErrorString(e);
        cudaGetLastError();  // a failed allocation is not sticky: clear it
        return e == cudaErrorMemoryAllocation ? DDM_ERR_NOMEM : DDM_ERR_CUDA;
    }
    return DDM_OK;
}

DDM_API int ddm_host_free(void *ptr) {
    if (!ptr) return DDM_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? DDM_OK : DDM_ERR_CUDA;
}
