// Issue cost of the MUFU flavours the simulator uses (cycles per warp instruction per scheduler), 16 warps per scheduler.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rates mufu_rates.cu && ./mufu_rates
#include <cstdio>
#include <cuda_runtime.h>
template <int W>
__global__ void __launch_bounds__(256) k(int iters, float seed, float *sink, unsigned long long *cyc) {
    float x[8];
#pragma unroll
    for (int c = 0; c < 8; c++) x[c] = seed + 0.001f * (threadIdx.x + c);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (W == 0) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (W == 1) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (W == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (W == 3) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (W == 4) asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (W == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (W == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (W == 7) { unsigned u = __float_as_uint(x[c]); asm volatile("popc.b32 %0, %0;" : "+r"(u)); x[c] = __uint_as_float(u | 0x3f800000u); }
            if (W == 8) asm volatile("{ .reg .f32 t; lg2.approx.ftz.f32 t, %0; abs.f32 t, t; sqrt.approx.ftz.f32 %0, t; }" : "+f"(x[c]));
        }
    }
    const long long t1 = clock64();
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < 8; c++) a += x[c];
    if (a == 1234.5f) sink[0] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}
template <int W>
double run(const char *name, int per) {
    float *sink; unsigned long long *cyc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8 * 148 * 8);
    const int iters = 4096, grid = 148 * 8;
    k<W><<<grid, 256>>>(16, 1.5f, sink, cyc);
    k<W><<<grid, 256>>>(iters, 1.5f, sink, cyc);
    cudaDeviceSynchronize();
    unsigned long long h[148 * 8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < grid; i++) s += (double)h[i];
    // 8 blocks x 8 warps per SM = 16 warps per scheduler, each issuing iters*8*per instructions in ~s/grid cycles
    const double cycles = s / grid, inst_per_sched = 16.0 * iters * 8 * per;
    printf("%-22s %.2f cycles per warp instruction per scheduler\n", name, cycles / inst_per_sched);
    cudaFree(sink); cudaFree(cyc);
    return 0;
}
int main() {
    run<0>("MUFU.LG2", 1); run<1>("MUFU.SQRT (sqrt.approx)", 1); run<2>("MUFU.RSQ", 1); run<3>("sin.approx (FMUL+MUFU)", 1);
    run<4>("cos.approx (FMUL+MUFU)", 1); run<5>("MUFU.EX2", 1); run<6>("MUFU.RCP", 1); run<7>("POPC (+LOP3)", 1); run<8>("lg2+sqrt pair", 1);
    return 0;
}
