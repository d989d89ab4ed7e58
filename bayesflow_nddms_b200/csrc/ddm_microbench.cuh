// ddm_microbench.cuh -- pipe micro-benchmarks behind ddm_microbench() (include/ddm_b200.h).
#pragma once
#include <cuda_runtime.h>

namespace ddm {
// Runs micro-benchmark `which` (ddm_microbench_id) on every SM.  *inst_per_s = warp-level
// instructions of the measured kind per second, chip-wide (for PHILOX / NORMALS: Philox blocks
// per warp per second); *sm_hz = SM clock seen (clock64 ticks / event time).
cudaError_t run_microbench(int which, int iters, int sm_count, cudaStream_t s, double *inst_per_s, double *sm_hz);
}  // namespace ddm
