// Micro-benchmark: issue/throughput of FFMA, packed FFMA2 (fma.rn.f32x2) and MUFU mixes on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ffma2bench.cu -o tools/_pb/ffma2
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <int MODE> __global__ void k(float* out, int iters, float seed) {
    float a[8]; u64 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; float2 v = make_float2(a[i], a[i] + 1.f); p[i] = *reinterpret_cast<u64*>(&v); }
    const float m = 1.0000001f, c = 1e-9f;
    float2 mv = make_float2(m, m), cv = make_float2(c, c);
    const u64 m2 = *reinterpret_cast<u64*>(&mv), c2 = *reinterpret_cast<u64*>(&cv);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {          // 16 FFMA (8 chains x 2)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
        } else if (MODE == 1) {   // 16 FFMA2 = 32 fma
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, c2);
        } else if (MODE == 2) {   // 16 FFMA + 2 MUFU
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
                a[r] = rsq(a[r]);
            }
        } else if (MODE == 3) {   // 8 FFMA2 (= 16 fma) + 2 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, c2);
            a[0] = rsq(a[0]); a[1] = rsq(a[1]);
        } else if (MODE == 4) {   // 8 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = rsq(a[i]);
        } else if (MODE == 5) {   // 8 FFMA2 + 8 FFMA (can they dual-issue?)
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], m2, c2); a[i] = fmaf(a[i], m, c); }
        } else if (MODE == 6) {   // 8 FFMA2 + 8 integer adds (alu pipe) 
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], m2, c2); a[i] = __int_as_float(__float_as_int(a[i]) + 3); }
        } else if (MODE == 7) {   // 8 FFMA + 8 integer adds
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], m, c); p[i] += 0x100000003ull; }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&p[i]); s += a[i] + v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double inst_per_iter, double flop_per_iter) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4, threads = 512, iters = 20000;
    float* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 100, 1.f); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); }
    const double warps = (double)blocks * threads / 32, winst = warps * iters * inst_per_iter;
    printf("%-34s %.3f ms  %.1f G warp-inst/s (%.2f per SM-cycle @1.965GHz)  %.1f TFLOP/s\n", name, best, winst / best * 1e-6,
           winst / (best * 1e-3) / (sms * 1.965e9), (double)blocks * threads * iters * flop_per_iter / best * 1e-9);
    cudaFree(out);
}
int main() {
    run<0>("16 FFMA", 16, 32);
    run<1>("16 FFMA2", 16, 64);
    run<2>("16 FFMA + 2 MUFU", 18, 32);
    run<3>("8 FFMA2 + 2 MUFU", 10, 32);
    run<4>("8 MUFU", 8, 0);
    run<5>("8 FFMA2 + 8 FFMA", 16, 48);
    run<6>("8 FFMA2 + 8 IADD", 16, 32);
    run<7>("8 FFMA + 8 IADD(64b)", 16, 16);
    return 0;
}
