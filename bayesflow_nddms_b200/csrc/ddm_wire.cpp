// Host side of the compact wire format (ddm_wire.cuh), compiled by g++ (no device code): a few threads turn each
// chunk's (steps, choice[, draw]) records into the caller's float64 (or float32) pairs with non-temporal stores
// while the GPU simulates the next chunk.  No simulation happens here: the Euler loop, the random numbers and the boundary test are the
// kernel's; this is the output formatting of basic_ddm_dc.py:108-112 applied to the kernel's integers.
#include "ddm_wire.cuh"

#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace ddm {

namespace {

struct WirePair {  // the kernel's int2 {code, fp32 bits}
    int32_t x, y;
};

const bool g_avx2 = [] {
    __builtin_cpu_init();  // this initialiser may run before libgcc's own constructor when the library is dlopen'ed
    return __builtin_cpu_supports("avx2") != 0;
}();

struct Segment {  // trials [lo, hi) of the chunk, all of one dataset
    int64_t lo, hi;
    double tau;
};

// Four trials per iteration: int32 codes -> (n*dt + tau, choice) float64 pairs, two 32-byte streaming stores.
// target("avx2") does not enable FMA: the product and the sum round separately, like the kernel's.
__attribute__((target("avx2"))) int64_t decode_basic64_avx2(const int32_t *w, double *o, int64_t lo, int64_t hi, double dt,
                                                            double tau, double timeout_val) {
    const __m256d vdt = _mm256_set1_pd(dt), vtau = _mm256_set1_pd(tau), vto = _mm256_set1_pd(timeout_val);
    const __m256d zero = _mm256_setzero_pd();
    const __m128i three = _mm_set1_epi32(3), one = _mm_set1_epi32(1);
    int64_t i = lo;
    for (; i + 4 <= hi; i += 4) {
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(w + i));
        const __m256d rt = _mm256_add_pd(_mm256_mul_pd(_mm256_cvtepi32_pd(_mm_srli_epi32(c, 2)), vdt), vtau);
        __m256d ch = _mm256_cvtepi32_pd(_mm_sub_epi32(_mm_and_si128(c, three), one));
        ch = _mm256_blendv_pd(ch, vto, _mm256_cmp_pd(ch, zero, _CMP_EQ_OQ));
        const __m256d a = _mm256_unpacklo_pd(rt, ch);  // rt0 ch0 | rt2 ch2
        const __m256d b = _mm256_unpackhi_pd(rt, ch);  // rt1 ch1 | rt3 ch3
        _mm256_stream_pd(o + 2 * i, _mm256_permute2f128_pd(a, b, 0x20));
        _mm256_stream_pd(o + 2 * i + 4, _mm256_permute2f128_pd(a, b, 0x31));
    }
    return i;
}

// (rt, choice) rows, float64.  Same operations as trial_outputs<true> in ddm_kernels.cuh.
void decode_basic64(const int32_t *w, double *o, const Segment &s, double dt, double timeout_val, bool stream) {
    const double lut[3] = {-1.0, timeout_val, 1.0};
    int64_t i = s.lo;
    if (stream && g_avx2) {
        for (; i < s.hi && ((reinterpret_cast<uintptr_t>(o + 2 * i) & 31u) != 0u); i++) {  // reach a 32-byte boundary
            const uint32_t c = (uint32_t)w[i];
            o[2 * i] = (double)(int32_t)(c >> 2) * dt + s.tau;
            o[2 * i + 1] = lut[c & 3u];
        }
        i = decode_basic64_avx2(w, o, i, s.hi, dt, s.tau, timeout_val);
    } else if (stream) {
        const __m128d vdt = _mm_set1_pd(dt), vtau = _mm_set1_pd(s.tau);
        for (; i + 2 <= s.hi; i += 2) {
            const uint32_t c0 = (uint32_t)w[i], c1 = (uint32_t)w[i + 1];
            const __m128i n = _mm_set_epi32(0, 0, (int)(c1 >> 2), (int)(c0 >> 2));
            const __m128d rt = _mm_add_pd(_mm_mul_pd(_mm_cvtepi32_pd(n), vdt), vtau);
            const __m128d ch = _mm_set_pd(lut[c1 & 3u], lut[c0 & 3u]);
            _mm_stream_pd(o + 2 * i, _mm_unpacklo_pd(rt, ch));
            _mm_stream_pd(o + 2 * i + 2, _mm_unpackhi_pd(rt, ch));
        }
    }
    for (; i < s.hi; i++) {
        const uint32_t c = (uint32_t)w[i];
        const double rt = (double)(int32_t)(c >> 2) * dt;
        o[2 * i] = rt + s.tau;
        o[2 * i + 1] = lut[c & 3u];
    }
}

// Eight trials per iteration: int32 codes -> (float)(n*dt + tau), choice as float32 pairs, two 32-byte streaming
// stores.  The response time is formed in double (multiply, add, then one rounding to float), exactly like the
// kernel's float32 store path (store_pair<false> over trial_outputs<true>).
__attribute__((target("avx2"))) int64_t decode_basic32_avx2(const int32_t *w, float *o, int64_t lo, int64_t hi, double dt,
                                                            double tau, float timeout_val) {
    const __m256d vdt = _mm256_set1_pd(dt), vtau = _mm256_set1_pd(tau);
    const __m256 vto = _mm256_set1_ps(timeout_val), zero = _mm256_setzero_ps();
    const __m256i three = _mm256_set1_epi32(3), one = _mm256_set1_epi32(1);
    int64_t i = lo;
    for (; i + 8 <= hi; i += 8) {
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(w + i));
        const __m256i n = _mm256_srli_epi32(c, 2);
        const __m256d rt0 = _mm256_add_pd(_mm256_mul_pd(_mm256_cvtepi32_pd(_mm256_castsi256_si128(n)), vdt), vtau);
        const __m256d rt1 = _mm256_add_pd(_mm256_mul_pd(_mm256_cvtepi32_pd(_mm256_extracti128_si256(n, 1)), vdt), vtau);
        const __m256 rt = _mm256_set_m128(_mm256_cvtpd_ps(rt1), _mm256_cvtpd_ps(rt0));
        __m256 ch = _mm256_cvtepi32_ps(_mm256_sub_epi32(_mm256_and_si256(c, three), one));
        ch = _mm256_blendv_ps(ch, vto, _mm256_cmp_ps(ch, zero, _CMP_EQ_OQ));
        const __m256 a = _mm256_unpacklo_ps(rt, ch);  // rt0 ch0 rt1 ch1 | rt4 ch4 rt5 ch5
        const __m256 b = _mm256_unpackhi_ps(rt, ch);  // rt2 ch2 rt3 ch3 | rt6 ch6 rt7 ch7
        _mm256_stream_ps(o + 2 * i, _mm256_permute2f128_ps(a, b, 0x20));
        _mm256_stream_ps(o + 2 * i + 8, _mm256_permute2f128_ps(a, b, 0x31));
    }
    return i;
}

void decode_basic32(const int32_t *w, float *o, const Segment &s, double dt, double timeout_val, bool stream) {
    const float lut[3] = {-1.f, (float)timeout_val, 1.f};
    int64_t i = s.lo;
    if (stream && g_avx2) {
        for (; i < s.hi && ((reinterpret_cast<uintptr_t>(o + 2 * i) & 31u) != 0u); i++) {  // reach a 32-byte boundary
            const uint32_t c = (uint32_t)w[i];
            o[2 * i] = (float)((double)(int32_t)(c >> 2) * dt + s.tau);
            o[2 * i + 1] = lut[c & 3u];
        }
        i = decode_basic32_avx2(w, o, i, s.hi, dt, s.tau, (float)timeout_val);
    }
    for (; i < s.hi; i++) {
        const uint32_t c = (uint32_t)w[i];
        const double rt = (double)(int32_t)(c >> 2) * dt;
        o[2 * i] = (float)(rt + s.tau);
        o[2 * i + 1] = lut[c & 3u];
    }
}

// (signed rt, external measurement) rows.  Same operations as trial_outputs<false>.
inline double signed_rt(uint32_t c, double dt, double tau) {
    const double rt = (double)(int32_t)(c >> 2) * dt;
    const int choice = (int)(c & 3u) - 1;
    return choice > 0 ? tau + rt : (choice < 0 ? -tau - rt : 0.0);
}

// Four (code, fp32 bits) records per iteration -> (signed rt, external measurement) float64 pairs, two 32-byte streaming
// stores.  signed rt = +(tau + n*dt) / -(tau + n*dt) / 0 by choice: -tau - rt of trial_outputs<false> is the same double
// as -(tau + rt) (rounding is symmetric), and the blends keep a missing response at +0.0 whatever tau's sign.
__attribute__((target("avx2"))) int64_t decode_ext64_avx2(const WirePair *w, double *o, int64_t lo, int64_t hi, double dt, double tau) {
    const __m256d vdt = _mm256_set1_pd(dt), vtau = _mm256_set1_pd(tau), zero = _mm256_setzero_pd();
    const __m256d sign = _mm256_set1_pd(-0.0);
    const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 1, 3, 5, 7);  // codes to the low half, fp32 bits to the high half
    const __m128i three = _mm_set1_epi32(3), one = _mm_set1_epi32(1);
    int64_t i = lo;
    for (; i + 4 <= hi; i += 4) {
        const __m256i r = _mm256_permutevar8x32_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(w + i)), pick);
        const __m128i c = _mm256_castsi256_si128(r);
        const __m256d ext = _mm256_cvtps_pd(_mm_castsi128_ps(_mm256_extracti128_si256(r, 1)));
        const __m256d t = _mm256_add_pd(vtau, _mm256_mul_pd(_mm256_cvtepi32_pd(_mm_srli_epi32(c, 2)), vdt));
        const __m256d ch = _mm256_cvtepi32_pd(_mm_sub_epi32(_mm_and_si128(c, three), one));
        __m256d v = _mm256_blendv_pd(zero, _mm256_xor_pd(t, sign), _mm256_cmp_pd(ch, zero, _CMP_LT_OQ));
        v = _mm256_blendv_pd(v, t, _mm256_cmp_pd(ch, zero, _CMP_GT_OQ));
        const __m256d a = _mm256_unpacklo_pd(v, ext);  // v0 e0 | v2 e2
        const __m256d b = _mm256_unpackhi_pd(v, ext);  // v1 e1 | v3 e3
        _mm256_stream_pd(o + 2 * i, _mm256_permute2f128_pd(a, b, 0x20));
        _mm256_stream_pd(o + 2 * i + 4, _mm256_permute2f128_pd(a, b, 0x31));
    }
    return i;
}

// Eight records per iteration -> (float signed rt, external measurement) float32 pairs.
__attribute__((target("avx2"))) int64_t decode_ext32_avx2(const WirePair *w, float *o, int64_t lo, int64_t hi, double dt, double tau) {
    const __m256d vdt = _mm256_set1_pd(dt), vtau = _mm256_set1_pd(tau), zero = _mm256_setzero_pd();
    const __m256d sign = _mm256_set1_pd(-0.0);
    const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 1, 3, 5, 7);
    const __m128i three = _mm_set1_epi32(3), one = _mm_set1_epi32(1);
    int64_t i = lo;
    for (; i + 8 <= hi; i += 8) {
        __m128 rt[2];
        __m128i eb[2];
        for (int half = 0; half < 2; half++) {
            const __m256i r = _mm256_permutevar8x32_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(w + i + 4 * half)), pick);
            const __m128i c = _mm256_castsi256_si128(r);
            eb[half] = _mm256_extracti128_si256(r, 1);
            const __m256d t = _mm256_add_pd(vtau, _mm256_mul_pd(_mm256_cvtepi32_pd(_mm_srli_epi32(c, 2)), vdt));
            const __m256d ch = _mm256_cvtepi32_pd(_mm_sub_epi32(_mm_and_si128(c, three), one));
            __m256d v = _mm256_blendv_pd(zero, _mm256_xor_pd(t, sign), _mm256_cmp_pd(ch, zero, _CMP_LT_OQ));
            v = _mm256_blendv_pd(v, t, _mm256_cmp_pd(ch, zero, _CMP_GT_OQ));
            rt[half] = _mm256_cvtpd_ps(v);
        }
        const __m256 vrt = _mm256_set_m128(rt[1], rt[0]);
        const __m256 vext = _mm256_castsi256_ps(_mm256_set_m128i(eb[1], eb[0]));
        const __m256 a = _mm256_unpacklo_ps(vrt, vext);  // r0 e0 r1 e1 | r4 e4 r5 e5
        const __m256 b = _mm256_unpackhi_ps(vrt, vext);  // r2 e2 r3 e3 | r6 e6 r7 e7
        _mm256_stream_ps(o + 2 * i, _mm256_permute2f128_ps(a, b, 0x20));
        _mm256_stream_ps(o + 2 * i + 8, _mm256_permute2f128_ps(a, b, 0x31));
    }
    return i;
}

void decode_ext64(const WirePair *w, double *o, const Segment &s, double dt, bool stream) {
    if (stream && g_avx2) {
        int64_t i = s.lo;
        for (; i < s.hi && ((reinterpret_cast<uintptr_t>(o + 2 * i) & 31u) != 0u); i++) {  // reach a 32-byte boundary
            float ext;
            __builtin_memcpy(&ext, &w[i].y, sizeof(ext));
            o[2 * i] = signed_rt((uint32_t)w[i].x, dt, s.tau);
            o[2 * i + 1] = (double)ext;
        }
        i = decode_ext64_avx2(w, o, i, s.hi, dt, s.tau);
        for (; i < s.hi; i++) {
            float ext;
            __builtin_memcpy(&ext, &w[i].y, sizeof(ext));
            o[2 * i] = signed_rt((uint32_t)w[i].x, dt, s.tau);
            o[2 * i + 1] = (double)ext;
        }
        return;
    }
    for (int64_t i = s.lo; i < s.hi; i++) {
        const WirePair r = w[i];
        float ext;
        static_assert(sizeof(ext) == sizeof(r.y), "fp32 bits");
        __builtin_memcpy(&ext, &r.y, sizeof(ext));
        const double o0 = signed_rt((uint32_t)r.x, dt, s.tau), o1 = (double)ext;
        if (stream) {
            _mm_stream_pd(o + 2 * i, _mm_set_pd(o1, o0));
        } else {
            o[2 * i] = o0;
            o[2 * i + 1] = o1;
        }
    }
}

void decode_ext32(const WirePair *w, float *o, const Segment &s, double dt, bool stream) {
    int64_t i0 = s.lo;
    if (stream && g_avx2) {
        for (; i0 < s.hi && ((reinterpret_cast<uintptr_t>(o + 2 * i0) & 31u) != 0u); i0++) {  // reach a 32-byte boundary
            float ext;
            __builtin_memcpy(&ext, &w[i0].y, sizeof(ext));
            o[2 * i0] = (float)signed_rt((uint32_t)w[i0].x, dt, s.tau);
            o[2 * i0 + 1] = ext;
        }
        i0 = decode_ext32_avx2(w, o, i0, s.hi, dt, s.tau);
    }
    for (int64_t i = i0; i < s.hi; i++) {
        const WirePair r = w[i];
        float ext;
        __builtin_memcpy(&ext, &r.y, sizeof(ext));
        o[2 * i] = (float)signed_rt((uint32_t)r.x, dt, s.tau);
        o[2 * i + 1] = ext;
    }
}

void decode_slice(const WireDecode &j, int id, int n_threads) {
    const int64_t total = j.n_datasets * j.n_trials;
    if (total == 0) return;
    // contiguous slices, even boundaries so that the 2-trial vector loop stays aligned to the slice start
    int64_t per = (total + n_threads - 1) / n_threads;
    per = (per + 1) & ~int64_t(1);
    const int64_t lo = std::min<int64_t>(total, per * id), hi = std::min<int64_t>(total, lo + per);
    if (lo >= hi) return;
    const bool stream = (reinterpret_cast<uintptr_t>(j.out) & (j.out64 ? 15u : 7u)) == 0u;  // rows are 16 / 8 bytes
    const double timeout_val = j.timeout_choice_one ? 1.0 : 0.0;
    int64_t ds = lo / j.n_trials;
    for (int64_t at = lo; at < hi; ds++) {
        Segment s;
        s.lo = at;
        s.hi = std::min<int64_t>(hi, (ds + 1) * j.n_trials);
        s.tau = j.params[(size_t)ds * j.n_params + j.tau_col];
        if (j.basic) {
            if (j.out64) decode_basic64(static_cast<const int32_t *>(j.wire), static_cast<double *>(j.out), s, j.dt, timeout_val, stream);
            else decode_basic32(static_cast<const int32_t *>(j.wire), static_cast<float *>(j.out), s, j.dt, timeout_val, stream);
        } else {
            if (j.out64) decode_ext64(static_cast<const WirePair *>(j.wire), static_cast<double *>(j.out), s, j.dt, stream && j.out64);
            else decode_ext32(static_cast<const WirePair *>(j.wire), static_cast<float *>(j.out), s, j.dt, stream);
        }
        at = s.hi;
    }
    if (stream) _mm_sfence();
}

}  // namespace

// Workers sleep on a condition variable between calls.  Inside a streamed call (begin() .. end()) they poll the
// job generation instead: waking 15 sleeping threads costs ~0.4 ms per chunk on the virtual machines measured
// (scripts/midsize_ab.py), as much as decoding a 2 Mi-trial chunk.
class HostWorkers {
  public:
    explicit HostWorkers(int n) : n_(n < 1 ? 1 : n) {
        for (int id = 1; id < n_; id++) threads_.emplace_back([this, id] { loop(id); });
    }
    ~HostWorkers() {
        stop_.store(true);
        kick();
        for (auto &t : threads_) t.join();
    }
    int size() const { return n_; }
    void begin() {
        hot_.store(true);
        kick();
    }
    void end() { hot_.store(false); }
    void run(HostSliceFn fn, const void *arg) {
        fn_ = fn;
        arg_ = arg;
        pending_.store(n_ - 1, std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        if (!hot_.load(std::memory_order_relaxed)) kick();
        fn(arg, 0, n_);
        while (pending_.load(std::memory_order_acquire) != 0) _mm_pause();
    }

  private:
    void kick() {
        { std::lock_guard<std::mutex> l(m_); }  // a waiter is either before its predicate check or already waiting
        wake_.notify_all();
    }
    void loop(int id) {
        uint64_t seen = 0;
        unsigned spins = 0;
        for (;;) {
            uint64_t g;
            while ((g = gen_.load(std::memory_order_acquire)) == seen) {
                if (stop_.load(std::memory_order_relaxed)) return;
                if (hot_.load(std::memory_order_relaxed)) {
                    // between the chunks of one streamed call: spin briefly (a chunk lands every few ms), then
                    // give the core away -- on a box with one rank per GPU the other ranks' threads want it
                    if (++spins < 4096) _mm_pause();
                    else std::this_thread::yield();
                    continue;
                }
                std::unique_lock<std::mutex> l(m_);
                wake_.wait(l, [&] { return stop_.load() || hot_.load() || gen_.load() != seen; });
            }
            seen = g;
            spins = 0;
            fn_(arg_, id, n_);
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    const int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable wake_;
    HostSliceFn fn_ = nullptr;
    const void *arg_ = nullptr;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0};
    std::atomic<bool> hot_{false}, stop_{false};
};

HostWorkers *host_workers_create(int n_threads) { return new HostWorkers(n_threads); }
void host_workers_destroy(HostWorkers *w) { delete w; }
int host_workers_size(const HostWorkers *w) { return w->size(); }
void host_workers_begin(HostWorkers *w) { w->begin(); }
void host_workers_end(HostWorkers *w) { w->end(); }

// The CPUs this process may run on, shared between the GPUs of the box (one process per GPU), at most 16:
// past that the decode is limited by host memory bandwidth, not by cores.
int host_workers_default_count(int gpus) {
    int cpus = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
    if (cpus <= 0) cpus = (int)std::thread::hardware_concurrency();
    if (gpus < 1) gpus = 1;
    const int n = cpus / gpus;
    return n < 1 ? 1 : (n > 16 ? 16 : n);
}

// The host-memory ceiling of the decode: n_threads threads fill `bytes` of `buf` with non-temporal 32-byte
// stores (the decode's own store instruction), nothing read.  Returns bytes per second.
namespace {
struct FillJob {
    char *buf;
    size_t bytes;
};
__attribute__((target("avx2"))) void fill_avx2(char *p, char *end) {
    const __m256d v = _mm256_set1_pd(1.0);
    for (; p + 32 <= end; p += 32) _mm256_stream_pd(reinterpret_cast<double *>(p), v);
}
void fill_slice(const void *arg, int id, int n) {
    const FillJob &j = *static_cast<const FillJob *>(arg);
    size_t per = (j.bytes / (size_t)n) & ~size_t(63);
    char *lo = j.buf + per * (size_t)id, *hi = (id == n - 1) ? j.buf + (j.bytes & ~size_t(63)) : lo + per;
    if (g_avx2) {
        fill_avx2(lo, hi);
    } else {
        const __m128d v = _mm_set1_pd(1.0);
        for (char *p = lo; p + 16 <= hi; p += 16) _mm_stream_pd(reinterpret_cast<double *>(p), v);
    }
    _mm_sfence();
}
}  // namespace

double host_stream_store_rate(HostWorkers *w, void *buf, size_t bytes, int reps) {
    FillJob j{static_cast<char *>(buf), bytes};
    w->run(fill_slice, &j);  // touch the pages
    double best = 0.0;
    for (int r = 0; r < reps; r++) {
        const auto t0 = std::chrono::steady_clock::now();
        w->run(fill_slice, &j);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt > 0.0 && (double)bytes / dt > best) best = (double)bytes / dt;
    }
    return best;
}

void wire_decode(HostWorkers *w, const WireDecode &job) {
    w->run([](const void *arg, int id, int n) { decode_slice(*static_cast<const WireDecode *>(arg), id, n); }, &job);
}
void host_workers_run(HostWorkers *w, HostSliceFn fn, const void *arg) { w->run(fn, arg); }

}  // namespace ddm
