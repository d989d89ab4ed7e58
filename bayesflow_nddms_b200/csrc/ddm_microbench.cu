// ddm_microbench.cu -- measures the per-pipe issue rates the DDM kernel's roofline is quoted
// against ("measure them on the box with micro-benchmarks before claiming
// a roofline").  Every kernel runs 8 blocks x 256 threads per SM (64 resident warps, the
// simulator's own occupancy class) with 8 independent dependency chains per thread.
#include <cstdint>
#include <cstdlib>

#include "../../include/ddm_b200.h"
#include "ddm_kernels.cuh"
#include "ddm_microbench.cuh"

namespace ddm {

constexpr int MB_CHAINS = 8;
constexpr int MB_UNROLL = 4;  // body = MB_CHAINS * MB_UNROLL measured instructions

struct MbOut {
    unsigned long long *cycles;  // per block
    uint32_t *sink;
};

template <int WHICH>
__global__ void __launch_bounds__(256) microbench_kernel(int iters, uint32_t seed, float fa, float fb, MbOut o,
                                                         const PhiloxKey key) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t xi[MB_CHAINS];
    float xf[MB_CHAINS];
    uint64_t xl[MB_CHAINS];
#pragma unroll
    for (int c = 0; c < MB_CHAINS; c++) {
        xi[c] = tid * 2654435761u + seed + c;
        xf[c] = 1.0f + (float)((tid + c) & 1023) * 9.765625e-4f;
        xl[c] = ((uint64_t)xi[c] << 32) | (xi[c] ^ 0x9e3779b9u);
    }
    const uint32_t ka = seed | 1u, kb = seed * 31u + 7u;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < MB_UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < MB_CHAINS; c++) {
                if (WHICH == DDM_MB_FFMA) {
                    xf[c] = __fmaf_rn(xf[c], fa, fb);
                } else if (WHICH == DDM_MB_IMAD_WIDE) {
                    xl[c] = (uint64_t)(uint32_t)xl[c] * (uint64_t)PHILOX_M0 + xl[c];
                } else if (WHICH == DDM_MB_LOP3) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(xi[c]) : "r"(ka), "r"(kb));
                } else if (WHICH == DDM_MB_IADD3) {
                    asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(xi[c]) : "r"(ka), "r"(kb));
                } else if (WHICH == DDM_MB_MUFU_LG2) {
                    asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(xf[c]));
                } else if (WHICH == DDM_MB_MUFU_SIN) {
                    asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(xf[c]));
                } else if (WHICH == DDM_MB_MIX_FMA_ALU) {
                    if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(xi[c]) : "r"(ka), "r"(kb));
                    else xf[c] = __fmaf_rn(xf[c], fa, fb);
                } else if (WHICH == DDM_MB_MIX_IMADW_LOP3) {  // 1 : 1, as inside a Philox round
                    if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(xi[c]) : "r"(ka), "r"(kb));
                    else xl[c] = (uint64_t)(uint32_t)xl[c] * (uint64_t)PHILOX_M0 + xl[c];
                } else if (WHICH == DDM_MB_MIX_MUFU_LOP3) {   // 1 : 3
                    if ((c & 3) == 0) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(xf[c]));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(xi[c]) : "r"(ka), "r"(kb));
                } else if (WHICH == DDM_MB_MIX_MUFU_IMADW) {  // 1 : 1
                    if (c & 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(xf[c]));
                    else xl[c] = (uint64_t)(uint32_t)xl[c] * (uint64_t)PHILOX_M0 + xl[c];
                } else if (WHICH == DDM_MB_MIX_BLOCKLIKE) {   // per 8: 1 MUFU, 2 IMAD.WIDE, 3 LOP3, 2 FFMA (the simulator block's mix)
                    if (c == 0) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(xf[c]));
                    else if (c == 1 || c == 2) xl[c] = (uint64_t)(uint32_t)xl[c] * (uint64_t)PHILOX_M0 + xl[c];
                    else if (c <= 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(xi[c]) : "r"(ka), "r"(kb));
                    else xf[c] = __fmaf_rn(xf[c], fa, fb);
                } else if (WHICH == DDM_MB_FFMA_REG) {        // three register operands (the Euler step's form)
                    xf[c] = __fmaf_rn(xf[c], xf[(c + 3) & 7], xf[(c + 5) & 7]);
                } else if (WHICH == DDM_MB_FADD_REG) {
                    xf[c] = __fadd_rn(xf[c], xf[(c + 3) & 7]);
                } else if (WHICH == DDM_MB_IMAD_WIDE_NOACC) {  // the Philox round's form: 32 x 32 -> 64, no addend
                    xl[c] = (uint64_t)((uint32_t)xl[c] ^ (uint32_t)(xl[c] >> 32)) * (uint64_t)PHILOX_M0;
                } else if (WHICH == DDM_MB_IMAD_HI) {
                    xi[c] = __umulhi(xi[c], PHILOX_M0) + ka;
                } else if (WHICH == DDM_MB_FFMA2) {            // packed fp32 pair (sm_100 fma.rn.f32x2), counted as one instruction
                    asm volatile("{ .reg .b64 a, b, d; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; fma.rn.f32x2 d, a, b, b; mov.b64 {%0, %1}, d; }"
                                 : "+f"(xf[c]), "+f"(xf[(c + 4) & 7]) : "f"(fa), "f"(fb));
                } else if (WHICH == DDM_MB_FSETP) {
                    asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(xi[c]) : "f"(xf[c]), "f"(fa));
                }
            }
        }
        if (WHICH == DDM_MB_PHILOX) {
            uint32_t w[4];
            philox4x32<10>((uint32_t)it, tid, xi[0], 0u, ka, kb, w);
            xi[0] ^= w[0] ^ w[1] ^ w[2] ^ w[3];
        } else if (WHICH == DDM_MB_PHILOX7) {
            uint32_t w[4];
            philox4x32<7>((uint32_t)it, tid, xi[0], 0u, ka, kb, w);
            xi[0] ^= w[0] ^ w[1] ^ w[2] ^ w[3];
        } else if (WHICH == DDM_MB_NORMALS2) {
            // two independent trials per lane: does instruction-level parallelism inside a warp buy issue slots?
            TrialF32 t;
            t.h = 3.4e38f; t.c0 = fb; t.u = fa; t.x = 0.f; t.ext = 0.f;
            uint32_t n = 0u, p = 1u, n2 = 0u, p2 = 1u;
            step_block_f32<false>((uint32_t)it, tid, xi[0], key, t, xf[0], n, p, 0xffffffffu);
            step_block_f32<false>((uint32_t)it, tid + 0x40000000u, xi[0], key, t, xf[1], n2, p2, 0xffffffffu);
            acc += n + n2;
        } else if (WHICH == DDM_MB_NORMALS) {
            // the simulator's own inner block: Philox -> 6 scaled normals -> 6 predicated Euler steps
            TrialF32 t;
            t.h = 3.4e38f; t.c0 = fb; t.u = fa; t.x = 0.f; t.ext = 0.f;
            uint32_t n = 0u, p = 1u;
            step_block_f32<false>((uint32_t)it, tid, xi[0], key, t, xf[0], n, p, 0xffffffffu);
            acc += n;
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int c = 0; c < MB_CHAINS; c++) acc ^= xi[c] ^ __float_as_uint(xf[c]) ^ (uint32_t)xl[c] ^ (uint32_t)(xl[c] >> 32);
    if (acc == 0x12345678u) o.sink[0] = acc;  // keeps the chains alive
    if (threadIdx.x == 0) o.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// DDM_MB_BLOCKS_PER_SM (1..8, default 8) limits the resident blocks per SM through the dynamic shared memory
// request, to measure a mix at the occupancy the simulator kernels actually run at.
static int mb_blocks_per_sm() {
    const char *e = getenv("DDM_MB_BLOCKS_PER_SM");
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
}

template <int WHICH>
static cudaError_t launch_one(int grid, int iters, MbOut o, cudaStream_t s) {
    const int bps = mb_blocks_per_sm();
    size_t smem = 0;
    if (bps < 8) {
        smem = (size_t)(220 * 1024 / bps) & ~size_t(1023);
        cudaError_t e = cudaFuncSetAttribute(microbench_kernel<WHICH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    microbench_kernel<WHICH><<<grid, 256, smem, s>>>(iters, 12345u, 1.0000001f, 1e-9f, o, make_philox_key(12345u, 678u));
    return cudaGetLastError();
}

static cudaError_t launch_which(int which, int grid, int iters, MbOut o, cudaStream_t s) {
    switch (which) {
    case DDM_MB_FFMA: return launch_one<DDM_MB_FFMA>(grid, iters, o, s);
    case DDM_MB_IMAD_WIDE: return launch_one<DDM_MB_IMAD_WIDE>(grid, iters, o, s);
    case DDM_MB_LOP3: return launch_one<DDM_MB_LOP3>(grid, iters, o, s);
    case DDM_MB_IADD3: return launch_one<DDM_MB_IADD3>(grid, iters, o, s);
    case DDM_MB_MUFU_LG2: return launch_one<DDM_MB_MUFU_LG2>(grid, iters, o, s);
    case DDM_MB_MUFU_SIN: return launch_one<DDM_MB_MUFU_SIN>(grid, iters, o, s);
    case DDM_MB_MIX_FMA_ALU: return launch_one<DDM_MB_MIX_FMA_ALU>(grid, iters, o, s);
    case DDM_MB_FSETP: return launch_one<DDM_MB_FSETP>(grid, iters, o, s);
    case DDM_MB_PHILOX: return launch_one<DDM_MB_PHILOX>(grid, iters, o, s);
    case DDM_MB_NORMALS: return launch_one<DDM_MB_NORMALS>(grid, iters, o, s);
    case DDM_MB_MIX_IMADW_LOP3: return launch_one<DDM_MB_MIX_IMADW_LOP3>(grid, iters, o, s);
    case DDM_MB_MIX_MUFU_LOP3: return launch_one<DDM_MB_MIX_MUFU_LOP3>(grid, iters, o, s);
    case DDM_MB_MIX_MUFU_IMADW: return launch_one<DDM_MB_MIX_MUFU_IMADW>(grid, iters, o, s);
    case DDM_MB_MIX_BLOCKLIKE: return launch_one<DDM_MB_MIX_BLOCKLIKE>(grid, iters, o, s);
    case DDM_MB_FFMA_REG: return launch_one<DDM_MB_FFMA_REG>(grid, iters, o, s);
    case DDM_MB_FADD_REG: return launch_one<DDM_MB_FADD_REG>(grid, iters, o, s);
    case DDM_MB_IMAD_WIDE_NOACC: return launch_one<DDM_MB_IMAD_WIDE_NOACC>(grid, iters, o, s);
    case DDM_MB_IMAD_HI: return launch_one<DDM_MB_IMAD_HI>(grid, iters, o, s);
    case DDM_MB_FFMA2: return launch_one<DDM_MB_FFMA2>(grid, iters, o, s);
    case DDM_MB_PHILOX7: return launch_one<DDM_MB_PHILOX7>(grid, iters, o, s);
    case DDM_MB_NORMALS2: return launch_one<DDM_MB_NORMALS2>(grid, iters, o, s);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t run_microbench(int which, int iters, int sm_count, cudaStream_t s, double *inst_per_s, double *sm_hz) {
    const int grid = sm_count * mb_blocks_per_sm();
    MbOut o{};
    cudaError_t e;
    if ((e = cudaMalloc(&o.cycles, sizeof(unsigned long long) * grid)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&o.sink, sizeof(uint32_t))) != cudaSuccess) { cudaFree(o.cycles); return e; }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    unsigned long long *h = new unsigned long long[grid];
    double best_s = 1e30, best_cycles = 0;
    for (int rep = 0; rep < 4 && e == cudaSuccess; rep++) {  // rep 0 warms up
        cudaEventRecord(e0, s);
        e = launch_which(which, grid, iters, o, s);
        cudaEventRecord(e1, s);
        if (e != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if ((e = cudaMemcpy(h, o.cycles, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        unsigned long long mx = 0;
        for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
        if (rep > 0 && ms * 1e-3 < best_s) { best_s = ms * 1e-3; best_cycles = (double)mx; }
    }
    delete[] h;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(o.cycles);
    cudaFree(o.sink);
    if (e != cudaSuccess) return e;
    const double warps = (double)grid * 256.0 / 32.0;
    const double per_iter = (which == DDM_MB_PHILOX || which == DDM_MB_NORMALS || which == DDM_MB_PHILOX7) ? 1.0
                          : which == DDM_MB_NORMALS2 ? 2.0
                          : (which == DDM_MB_MIX_FMA_ALU ? (double)(MB_CHAINS * MB_UNROLL) : (double)(MB_CHAINS * MB_UNROLL));
    *inst_per_s = warps * (double)iters * per_iter / best_s;
    // the kernel's longest block spans (almost) the whole launch: cycles / time = SM clock
    *sm_hz = best_cycles / best_s;
    return cudaSuccess;
}

}  // namespace ddm
