// Shared device helpers for the csf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include "csf_b200.h"

#define CSF_PI 3.14159265358979323846
#define CSF_TWO_PI 6.28318530717958647692

void csf_set_error(const char* where, cudaError_t e);
#define CSF_CHECK_LAUNCH(where)                         \
    do {                                                \
        cudaError_t e__ = cudaGetLastError();           \
        if (e__ != cudaSuccess) {                       \
            csf_set_error(where, e__);                  \
            return -(int)e__;                           \
        }                                               \
    } while (0)

// ---- pair payload element -------------------------------------------------------
template <typename T> struct Xycs;
template <> struct __align__(16) Xycs<float> { int32_t xq, yq; float c, s; };
template <> struct __align__(16) Xycs<double> { double x, y, c, s; };

// ---- scalar math traits -----------------------------------------------------------
// float: single MUFU approximations (rsqrt/sqrt/ex2/rcp .approx.ftz, rel. error <= 2^-22)
// double: IEEE functions.
template <typename T> struct M;
template <> struct M<float> {
    static __device__ __forceinline__ float rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
};
template <> struct M<double> {
    static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double ex2(double x) { return ::exp2(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
};

// Division and square root of the per-agent kernels: IEEE in the f64 verification build; in the f32
// production build the single-instruction MUFU approximations (relative error <= 2^-22, far inside the
// 1e-4 per-step budget) -- an IEEE float division is ~10 instructions with a slow path, and the per-agent
// kernel is bound by instruction fetch and issue latency, not by bandwidth.
__device__ __forceinline__ float qdiv(float a, float b) { return a * M<float>::rcp(b); }
__device__ __forceinline__ double qdiv(double a, double b) { return a / b; }
__device__ __forceinline__ float qsqrt(float a) { return M<float>::sqrt(a); }
__device__ __forceinline__ double qsqrt(double a) { return ::sqrt(a); }

// ---- angles (reference utils.py) --------------------------------------------------
// limitAngle, utils.py:124-139: wrap to (-pi, pi]
__device__ __forceinline__ float periods(float th) { return floorf(th * (float)(1.0 / CSF_TWO_PI)); }
__device__ __forceinline__ double periods(double th) { return floor(th * (1.0 / CSF_TWO_PI)); }   // (the fold below absorbs a last-bit difference at multiples of 2 pi)
template <typename T> __device__ __forceinline__ T limit_angle(T th) {
    const T tp = (T)CSF_TWO_PI, pi = (T)CSF_PI;
    th = periods(th) * (-tp) + th;
    if (th > pi) th -= tp;
    else if (th < -pi) th += tp;
    return th;
}
// angleDifference, utils.py:151-182: signed shortest rotation a1 -> a2 (ties -> +)
template <typename T> __device__ __forceinline__ T angle_difference(T a1, T a2) {
    const T tp = (T)CSF_TWO_PI, pi = (T)CSF_PI;
    T da = (a1 > a2) ? (a1 - a2) : (a2 - a1);
    if (da > pi) da = tp - da;
    T t1 = fabs(limit_angle(a1 - da) - a2);
    T t2 = fabs(limit_angle(a1 + da) - a2);
    return (t1 < t2) ? -da : da;
}
template <typename T> __device__ __forceinline__ T clampT(T x, T lo, T hi) {  // utils.thresh :204-227
    return fmax(fmin(x, hi), lo);
}
__device__ __forceinline__ void sincosT(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincosT(double a, double* s, double* c) { sincos(a, s, c); }

// Q-format quantisation of a position for the f32 payload.
__device__ __forceinline__ int32_t quantise(double x, double inv_q, int* overflow) {
    double r = rint(x * inv_q);
    if (fabs(r) > 1073741824.0) { *overflow = 1; r = fmax(fmin(r, 1073741824.0), -1073741824.0); }
    return (int32_t)r;
}
// (ox, oy) = origin of the Q-format frame: centring the frame on the crowd spends the 31 bits on the
// occupied region only (the pair kernel sees differences, the origin drops out)
template <typename T> __device__ __forceinline__ Xycs<T> store_xycs(void* out, int64_t idx, double x, double y, T psi,
                                                                    double inv_q, double ox, double oy, int* overflow);
template <> __device__ __forceinline__ Xycs<float> store_xycs<float>(void* out, int64_t idx, double x, double y, float psi,
                                                                     double inv_q, double ox, double oy, int* overflow) {
    Xycs<float> e;
    e.xq = quantise(x - ox, inv_q, overflow);
    e.yq = quantise(y - oy, inv_q, overflow);
    sincosf(psi, &e.s, &e.c);
    reinterpret_cast<Xycs<float>*>(out)[idx] = e;
    return e;
}
template <> __device__ __forceinline__ Xycs<double> store_xycs<double>(void* out, int64_t idx, double x, double y, double psi,
                                                                       double, double, double, int*) {
    Xycs<double> e;
    e.x = x; e.y = y;
    sincos(psi, &e.s, &e.c);
    reinterpret_cast<Xycs<double>*>(out)[idx] = e;
    return e;
}
