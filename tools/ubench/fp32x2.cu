// Micro-benchmark: issue/pipe throughput of scalar FP32 vs packed f32x2 math on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void k(float* out, float s) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float b = s, c = s * 0.5f;
    unsigned long long pa0, pa1, pa2, pa3, pb, pc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(pa0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pa1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pa2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pa3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pc) : "f"(c), "f"(c));
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
        if (MODE == 0) {  // 8 scalar FFMA
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        } else if (MODE == 1) {  // 4 packed FFMA2 (same flops as mode 0)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa0) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa1) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa2) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa3) : "l"(pb), "l"(pc));
        } else if (MODE == 2) {  // 8 scalar FADD
            a0 += b; a1 += b; a2 += b; a3 += b; a4 += b; a5 += b; a6 += b; a7 += b;
        } else if (MODE == 3) {  // 4 packed FADD2
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa0) : "l"(pb));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa1) : "l"(pb));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa2) : "l"(pb));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa3) : "l"(pb));
        } else if (MODE == 4) {  // 4 FFMA2 + 4 scalar IADD (issue mix)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa0) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa1) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa2) : "l"(pb), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa3) : "l"(pb), "l"(pc));
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        } else if (MODE == 5) {  // 8 scalar FMUL
            a0 *= b; a1 *= b; a2 *= b; a3 *= b; a4 *= b; a5 *= b; a6 *= b; a7 *= b;
        }
    }
    long long t1 = clock64();
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(pa0));
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(pa1)); r += a0 + a1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(pa2)); r += a0 + a1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(pa3)); r += a0 + a1;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[1 << 20] = t1 - t0;
}
template <int MODE> void run(const char* name, float* d, int warps_per_sm, int flop_lanes_per_iter) {
    int threads = warps_per_sm * 32;
    k<MODE><<<148, threads>>>(d, 1.0001f);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(d, 1.0001f);
    cudaDeviceSynchronize();
    long long cyc; cudaMemcpy(&cyc, ((long long*)d) + (1 << 20), 8, cudaMemcpyDeviceToHost);
    double per_iter = (double)cyc / ITER;
    printf("%-28s warps/SM %2d  cycles/iter %.2f  -> %.1f fp32 lane-results/clk/SM (warp-instr/clk/SMSP %.3f)\n", name, warps_per_sm,
           per_iter, warps_per_sm * 32.0 * flop_lanes_per_iter / per_iter, 0.0);
}
int main() {
    float* d; cudaMalloc(&d, (1 << 23) + 64);
    for (int w : {4, 8, 16, 32}) {
        run<0>("8x FFMA", d, w, 8);
        run<1>("4x FFMA2 (f32x2)", d, w, 8);
        run<2>("8x FADD", d, w, 8);
        run<3>("4x FADD2 (f32x2)", d, w, 8);
        run<4>("4x FFMA2 + 4x FFMA", d, w, 12);
        run<5>("8x FMUL", d, w, 8);
    }
    return 0;
}
