// Compact device->host wire format of ddm_simulate's chunked pipeline, and its host-side decoder.
//
// A finished trial is (Euler steps n, choice in {-1, 0, +1}) plus, for the models with a second measured
// column, one fp32 draw.  The reference's output values are a fixed function of those and of the dataset's
// non-decision time (basic_ddm_dc.py:108-112: rt = n*dt + ndt), so the stepping kernel can ship 4 (or 8)
// bytes per trial over PCIe instead of 16, and the host writes the float64 pairs the caller asked for while
// the next chunk is still being simulated.  The decode performs the same two IEEE double operations
// (multiply, then add: no contraction) the kernel's own float64 store path performs, so the two paths give
// identical bits (tests/test_gpu_parity.py::test_compact_wire_matches_plain_copy).
#pragma once
#include <cstddef>
#include <cstdint>

namespace ddm {

// word 0 of a wire record: (n << 2) | (choice + 1); n < 2^29
constexpr uint32_t WIRE_MAX_STEPS = (1u << 29) - 1u;

#ifdef __CUDACC__
__host__ __device__
#endif
inline int32_t wire_pack(uint32_t n, int choice) { return (int32_t)((n << 2) | (uint32_t)(choice + 1)); }

struct WireDecode {
    const void *wire;       // chunk: int32 per trial (basic layout) or int2 {code, fp32 bits} per trial
    void *out;              // chunk destination: (rows, 2) float64 or float32
    const double *params;   // host parameters of the chunk's first dataset
    int n_params, tau_col;
    int64_t n_datasets, n_trials;
    double dt;
    bool basic;             // (rt, choice) columns; else (signed rt, external measurement)
    bool out64;
    bool timeout_choice_one;
};

// A small pool of host threads (the caller takes part) that decode one chunk at a time.
class HostWorkers;
typedef void (*HostSliceFn)(const void *arg, int slice, int n_slices);  // runs once per thread, slice = 0 .. n_slices-1
HostWorkers *host_workers_create(int n_threads);  // n_threads >= 1 (1 = the calling thread only)
void host_workers_destroy(HostWorkers *w);
int host_workers_size(const HostWorkers *w);
// bracket a streamed call: in between, idle workers poll for the next chunk instead of sleeping
void host_workers_begin(HostWorkers *w);
void host_workers_end(HostWorkers *w);
int host_workers_default_count(int gpus_on_box);
void wire_decode(HostWorkers *w, const WireDecode &job);
void host_workers_run(HostWorkers *w, HostSliceFn fn, const void *arg);
// bytes per second the pool's threads reach filling `buf` with streaming stores (the decode's ceiling)
double host_stream_store_rate(HostWorkers *w, void *buf, size_t bytes, int reps);

}  // namespace ddm
