// Device code shared by the dense (csf_pair.cu) and the tiled/culled (csf_pair_tiled.cu) pair
// kernels: the per-pair force evaluation and the mbarrier / TMA bulk-copy primitives.
#pragma once
#include "csf_common.cuh"

namespace {

template <typename T> struct PairConst {
    T sg0, sg1, sg2, sg3;  // sigma_0..3 / (q_scale * log2 e)
    T e0, e1;
    T qa, qb, qc;          // 1 - e^2 = qa + qb s2 + qc s2^2  (qa = 1-e0^2, qb = 2 e0 e1, qc = -e1^2)
    T ncosH;               // -cos(hfov/2); +2 if hfov/2 >= pi (always visible)
    T tiny;                // guard added to rho^2 (coincident pair -> zero contribution)
};

template <typename T> struct Tgt;
template <> struct Tgt<float> { int32_t xq, yq; float c, s; };
template <> struct Tgt<double> { double x, y, c, s; };

__device__ __forceinline__ void delta(const Xycs<float>& sr, const Tgt<float>& tg, float& dx, float& dy) {
    dx = (float)(tg.xq - sr.xq);  // exact integer difference, then one rounding
    dy = (float)(tg.yq - sr.yq);
}
__device__ __forceinline__ void delta(const Xycs<double>& sr, const Tgt<double>& tg, double& dx, double& dy) {
    dx = tg.x - sr.x;
    dy = tg.y - sr.y;
}
__device__ __forceinline__ float mulsign(float v, float s) {  // v * sign(s) for s != 0 (sign bit xor)
    return __int_as_float(__float_as_int(v) ^ (__float_as_int(s) & 0x80000000));
}
__device__ __forceinline__ double mulsign(double v, double s) {
    return __longlong_as_double(__double_as_longlong(v) ^ (__double_as_longlong(s) & 0x8000000000000000ll));
}

template <typename T, bool P2R>
__device__ __forceinline__ void pair_eval(const Xycs<T>& sr, const Tgt<T>& tg, const PairConst<T>& k, T& ax, T& ay) {
    T dx, dy;
    delta(sr, tg, dx, dy);
    const T r2 = fma(dy, dy, fma(dx, dx, k.tiny));
    const T rinv = M<T>::rsqrt(r2);
    const T ux = dx * rinv, uy = dy * rinv;
    const T c = fma(uy, sr.s, ux * sr.c);
    const T s = fma(-ux, sr.s, uy * sr.c);
    const T t = fma(uy, tg.s, ux * tg.c);
    bool vis = t <= k.ncosH;
    if (P2R) vis = vis && (fma(tg.s, ux, -(tg.c * uy)) <= (T)0);
    const T sr_ = fma(sr.s, tg.c, -(sr.c * tg.s));
    const T s2 = sr_ * sr_;
    const T A = fma(k.sg1, s2, k.sg0);
    const T B = fma(k.sg3, s2, k.sg2);
    const T e = fma(-k.e1, s2, k.e0);
    // half angles |sin(phi/2)|, |cos(phi/2)| without the 1 -/+ c cancellation:
    //   big = sqrt((1+|c|)/2),  small = |s| / (2 big);   (h1, h2) = c >= 0 ? (small, big) : (big, small)
    const T hm = fma((T)0.5, fabs(c), (T)0.5);
    const T rm = M<T>::rsqrt(hm);
    const T hbig = hm * rm;
    const T hsmall = fabs(s) * ((T)0.5 * rm);
    const bool fwd = c >= (T)0;
    const T h1 = fwd ? hsmall : hbig;
    // sign(phi) = sign(s) is 0 for s == 0 (np.sign, vehicle.py:1625): there the sigma' term must
    // vanish; picking hsmall (= 0 when s == 0) for h2 does that without a separate select.
    const T h2 = (fwd && s != (T)0) ? hbig : hsmall;
    const T sg = fma(-B, h1, A);
    // q^2 = 1 - (e c)^2 = (1 - e^2) + (e s)^2: no cancellation when the target sits on the source's
    // axis (|c| -> 1, q^2 -> 1 - e0^2 = 0.01), where an error of 2^-22 in rinv would otherwise be
    // amplified 200x into the exponent.
    const T ec = e * c;
    const T es = e * fabs(s);
    const T q2 = fma(es, es, fma(fma(k.qc, s2, k.qb), s2, k.qa));
    const T w = es * ec;
    const T mv = fma(w, sg, (T)0.5 * (q2 * (B * h2)));
    const T grho = q2 * sg;
    const T gphi = mulsign(mv, s);  // mv == 0 when s == 0 (w = 0 and h2 = 0)
    const T ry = M<T>::rsqrt(grho * sg);
    const T qs = q2 * ry;
    const T rho = r2 * rinv;
    const T P = M<T>::ex2(-(rho * qs));
    const T rn = M<T>::rsqrt(fma(gphi, gphi, grho * grho));
    const T sc = vis ? P * rn : (T)0;
    const T a = sc * grho, b = sc * gphi;
    ax = fma(a, ux, ax);
    ax = fma(-b, uy, ax);
    ay = fma(a, uy, ay);
    ay = fma(b, ux, ay);
}

// ---- mbarrier / TMA bulk-copy primitives (sm_90+/sm_100a) --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T> inline PairConst<T> make_const(const CsfFieldParams* fp, bool is_f32) {
    const double kappa = is_f32 ? fp->q_scale * 1.4426950408889634 : 1.4426950408889634;
    PairConst<T> k;
    k.sg0 = (T)(fp->sigma_0 / kappa);
    k.sg1 = (T)(fp->sigma_1 / kappa);
    k.sg2 = (T)(fp->sigma_2 / kappa);
    k.sg3 = (T)(fp->sigma_3 / kappa);
    k.e0 = (T)fp->e_0;
    k.e1 = (T)fp->e_1;
    k.qa = (T)(1.0 - fp->e_0 * fp->e_0);
    k.qb = (T)(2.0 * fp->e_0 * fp->e_1);
    k.qc = (T)(-fp->e_1 * fp->e_1);
    k.ncosH = (fp->hfov * 0.5 >= CSF_PI) ? (T)2 : (T)(-cos(fp->hfov * 0.5));
    k.tiny = (T)(is_f32 ? 1e-6 : 1e-200);
    return k;
}


}  // namespace
