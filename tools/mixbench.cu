// Micro-benchmark: the pair kernel's instruction mix (per 4 pair evaluations of a lane: 96 packed FP32x2,
// 20 MUFU, ~46 ALU-pipe, ~10 IMAD) as INDEPENDENT chains -- the throughput the SM's pipes allow for this
// mix when nothing depends on anything -- and variations that isolate one pipe each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mixbench.cu -o tools/_pb/mix && tools/_pb/mix
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// NF2 packed FMAs (distinct 64-bit operands if DIST), NMU MUFU, NALU integer-ALU ops, per iteration, on 8 chains
template <int NF2, int NMU, int NALU, bool DIST> __global__ void k(float* out, int iters, float seed) {
    u64 p[8]; float a[8]; int q[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; float2 v = make_float2(a[i], a[i] + 1.f); p[i] = *reinterpret_cast<u64*>(&v); q[i] = threadIdx.x + i; }
    float2 mv = make_float2(1.0000001f, 0.9999999f), cv = make_float2(1e-9f, 2e-9f);
    const u64 m2 = *reinterpret_cast<u64*>(&mv), c2 = *reinterpret_cast<u64*>(&cv);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NF2; ++j) {
            const int i = j & 7;
            p[i] = DIST ? fma2(p[i], p[(i + 3) & 7], p[(i + 5) & 7]) : fma2(p[i], m2, c2);
        }
#pragma unroll
        for (int j = 0; j < NMU; ++j) a[j & 7] = rsq(a[j & 7]);
#pragma unroll
        for (int j = 0; j < NALU; ++j) { const int i = j & 7; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(it)); }
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&p[i]); s += a[i] + v.x + v.y + (float)q[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF2, int NMU, int NALU, bool DIST> void run(const char* name, int warps_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 32 * warps_per_sm / 2, blocks = sms * 2, iters = 4000;
    float* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NF2, NMU, NALU, DIST><<<blocks, threads>>>(out, 100, 1.f); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); k<NF2, NMU, NALU, DIST><<<blocks, threads>>>(out, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); }
    // cycles per iteration and scheduler: time * clock / (iters * warps per scheduler)
    const double cyc = best * 1e-3 * 1.965e9 / ((double)iters * warps_per_sm / 4.0);
    printf("%-52s %2d warps/SM  %.3f ms  %.1f cycles per (iteration, warp) and scheduler\n", name, warps_per_sm, best, cyc);
    cudaFree(out);
}
int main() {
    for (int w : {16, 24}) {
        if (w == 16) {
            run<96, 0, 0, false>("96 FFMA2 (shared operands)", 16);
            run<96, 0, 0, true>("96 FFMA2 (3 distinct 64-bit operands)", 16);
            run<0, 20, 0, false>("20 MUFU", 16);
            run<0, 0, 56, false>("56 LOP3", 16);
            run<96, 20, 0, true>("96 FFMA2 + 20 MUFU", 16);
            run<96, 0, 56, true>("96 FFMA2 + 56 LOP3", 16);
            run<96, 20, 56, true>("96 FFMA2 + 20 MUFU + 56 LOP3 (the kernel's mix)", 16);
        } else {
            run<96, 20, 56, true>("96 FFMA2 + 20 MUFU + 56 LOP3 (the kernel's mix)", 24);
        }
    }
    return 0;
}
