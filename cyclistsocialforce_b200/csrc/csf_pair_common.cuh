// Device code shared by the dense (csf_pair.cu) and the tiled/culled (csf_pair_tiled.cu) pair
// kernels: the per-pair force evaluation and the mbarrier / TMA bulk-copy primitives.
#pragma once
#include "csf_common.cuh"

namespace {

template <typename T> struct PairConst {
    T sg0, sg1, sg2, sg3;  // sigma_0..3 / (q_scale * log2 e)
    T sg2s, sg3s;          // sg2 / sqrt 2, sg3 / sqrt 2 (pair_eval2 works with sqrt 2 x the half-angle functions)
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

// ---- packed FP32x2 evaluation (Blackwell FFMA2) ------------------------------------------------------
// Two sources against one target per lane, every FP32 add/mul/fma as one fma.rn.f32x2 (SASS FFMA2:
// same FP32-pipe throughput as two FFMA, half the issue slots -- the kernel is issue bound).  The
// operation sequence per half is exactly pair_eval<float>'s.  neg/abs of both halves and scalar
// broadcasts are written as plain moves; ptxas folds them into FFMA2 operand modifiers.
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 pk(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up(F2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 splat(float s) { return pk(s, s); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 neg2(F2 a) { float l, h; up(a, l, h); return pk(-l, -h); }
__device__ __forceinline__ F2 abs2(F2 a) { float l, h; up(a, l, h); return pk(fabsf(l), fabsf(h)); }

// sources (x0,y0,c0,s0) and (x1,y1,c1,s1); accumulators hold one partial sum per half
template <bool P2R>
__device__ __forceinline__ void pair_eval2(int32_t x0, int32_t x1, int32_t y0, int32_t y1, float c0, float c1, float s0,
                                           float s1, const Tgt<float>& tg, const PairConst<float>& k, F2& ax, F2& ay) {
    const F2 DX = pk((float)(tg.xq - x0), (float)(tg.xq - x1));
    const F2 DY = pk((float)(tg.yq - y0), (float)(tg.yq - y1));
    const F2 R2 = fma2(DY, DY, fma2(DX, DX, splat(k.tiny)));
    float r2a, r2b;
    up(R2, r2a, r2b);
    const F2 RINV = pk(M<float>::rsqrt(r2a), M<float>::rsqrt(r2b));
    const F2 UX = mul2(DX, RINV), UY = mul2(DY, RINV);
    const F2 SC = pk(c0, c1), SS = pk(s0, s1), TC = splat(tg.c), TS = splat(tg.s);
    const F2 C = fma2(UY, SS, mul2(UX, SC));
    const F2 S = fma2(neg2(UX), SS, mul2(UY, SC));
    const F2 TT = fma2(UY, TS, mul2(UX, TC));
    float ta, tb;
    up(TT, ta, tb);
    bool visa = ta <= k.ncosH, visb = tb <= k.ncosH;
    if (P2R) {
        const F2 L = fma2(TS, UX, neg2(mul2(TC, UY)));
        float la, lb;
        up(L, la, lb);
        visa = visa && (la <= 0.f);
        visb = visb && (lb <= 0.f);
    }
    const F2 SR = fma2(SS, TC, neg2(mul2(SC, TS)));
    const F2 S2 = mul2(SR, SR);
    const F2 A = fma2(splat(k.sg1), S2, splat(k.sg0));
    // half-angle functions of phi scaled by sqrt 2: with hm2 = 1 + |c| = 2 hm,  sqrt(2 hm) = hm2 rsqrt(hm2)  and
    // |s| / sqrt(2 hm) = |s| rsqrt(hm2) -- one packed instruction less than via hm = (1 + |c|) / 2; the factor
    // 1 / sqrt 2 sits in the constants of B (B only ever multiplies one of the two)
    const F2 B = fma2(splat(k.sg3s), S2, splat(k.sg2s));
    const F2 E = fma2(splat(-k.e1), S2, splat(k.e0));
    const F2 AS = abs2(S);
    const F2 HM = add2(abs2(C), splat(1.f));
    float hma, hmb;
    up(HM, hma, hmb);
    const F2 RM = pk(M<float>::rsqrt(hma), M<float>::rsqrt(hmb));
    const F2 HBIG = mul2(HM, RM);
    const F2 HSMALL = mul2(AS, RM);
    float ca, cb, sa, sb, hba, hbb, hsa, hsb;
    up(C, ca, cb);
    up(S, sa, sb);
    up(HBIG, hba, hbb);
    up(HSMALL, hsa, hsb);
    const bool fwa = ca >= 0.f, fwb = cb >= 0.f;
    const F2 H1 = pk(fwa ? hsa : hba, fwb ? hsb : hbb);
    const F2 H2 = pk((fwa && sa != 0.f) ? hba : hsa, (fwb && sb != 0.f) ? hbb : hsb);
    const F2 SG = fma2(neg2(B), H1, A);
    const F2 EC = mul2(E, C);
    const F2 ES = mul2(E, AS);
    const F2 Q2 = fma2(ES, ES, fma2(fma2(splat(k.qc), S2, splat(k.qb)), S2, splat(k.qa)));
    const F2 W = mul2(ES, EC);
    const F2 MV = fma2(W, SG, mul2(splat(0.5f), mul2(Q2, mul2(B, H2))));
    const F2 GRHO = mul2(Q2, SG);
    float mva, mvb;
    up(MV, mva, mvb);
    const F2 GPHI = pk(mulsign(mva, sa), mulsign(mvb, sb));
    const F2 RYA = mul2(GRHO, SG);
    float rya, ryb;
    up(RYA, rya, ryb);
    const F2 RY = pk(M<float>::rsqrt(rya), M<float>::rsqrt(ryb));
    const F2 QS = mul2(Q2, RY);
    const F2 RHO = mul2(R2, RINV);
    const F2 PE = mul2(RHO, QS);
    float pea, peb;
    up(PE, pea, peb);
    const F2 P = pk(M<float>::ex2(-pea), M<float>::ex2(-peb));
    const F2 RNA = fma2(GPHI, GPHI, mul2(GRHO, GRHO));
    float rna, rnb;
    up(RNA, rna, rnb);
    const F2 RN = pk(M<float>::rsqrt(rna), M<float>::rsqrt(rnb));
    const F2 SCL = mul2(P, RN);
    float sca, scb;
    up(SCL, sca, scb);
    const F2 SCV = pk(visa ? sca : 0.f, visb ? scb : 0.f);
    const F2 FA = mul2(SCV, GRHO), FB = mul2(SCV, GPHI);
    ax = fma2(FA, UX, ax);
    ax = fma2(neg2(FB), UY, ax);
    ay = fma2(FA, UY, ay);
    ay = fma2(FB, UX, ay);
}

// ---- Bicycle v0.1 elliptic field (reference vehicle.py:1054-1147) for the tiled kernel ---------------
// e = min((v / v_max)^0.1, 0.7) (updateExcentricity :1054-1064);  b = rho (1 - e cos phi0) / (sqrt(1-e^2) p_decay);
// P = p_0 exp(-b) / p_decay;  F_rho = P (1 - e cos phi0) / sqrt(1-e^2);  F_phi = P e sin phi0 / sqrt(1-e^2).
// The sorted copy of such sources carries the heading scaled by the eccentricity, (e cos psi, e sin psi): the
// field needs nothing else of the source -- e cos phi0 and e sin phi0 are the two products with the unit vector
// towards the target, e^2 is the squared length -- so a source stays one 16-byte (f64: 32-byte) element and the
// tile layout, the stages and the survivor buffers are those of the TwoD field.  k.sg0 = log2(e) / p_decay in
// payload units; the sum is scaled by p_0 / p_decay where the partial sums are reduced.
template <typename T, bool P2R>
__device__ __forceinline__ void pair_eval_bike(const Xycs<T>& sr, const Tgt<T>& tg, const PairConst<T>& k, T& ax, T& ay) {
    T dx, dy;
    delta(sr, tg, dx, dy);
    const T r2 = fma(dy, dy, fma(dx, dx, k.tiny));
    const T rinv = M<T>::rsqrt(r2);
    const T ux = dx * rinv, uy = dy * rinv;
    const T ec = fma(uy, sr.s, ux * sr.c);             // e cos phi0
    const T es = fma(-ux, sr.s, uy * sr.c);            // e sin phi0
    const T t = fma(uy, tg.s, ux * tg.c);
    bool vis = t <= k.ncosH;
    if (P2R) vis = vis && (fma(tg.s, ux, -(tg.c * uy)) <= (T)0);
    const T ke = M<T>::rsqrt(fma(-sr.s, sr.s, fma(-sr.c, sr.c, (T)1)));   // 1 / sqrt(1 - e^2)
    const T g = ((T)1 - ec) * ke;
    const T rho = r2 * rinv;
    const T P = M<T>::ex2(-(rho * g * k.sg0));
    const T fr = vis ? P * g : (T)0;
    const T fp = vis ? P * (es * ke) : (T)0;
    ax = fma(fr, ux, ax);
    ax = fma(-fp, uy, ax);
    ay = fma(fr, uy, ay);
    ay = fma(fp, ux, ay);
}

// ---- mbarrier / TMA bulk-copy primitives (sm_90+/sm_100a) --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// adds to the transaction count of the current phase without arriving (several batches of copies
// may be announced before the one arrival that lets the phase complete)
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// producer-side wait: the producer is usually far ahead of the consumers; sleep between polls so
// that its spinning does not take issue slots from the consumer warps of its scheduler
#ifndef CSF_WAIT_HINT_NS
#define CSF_WAIT_HINT_NS 20000u
#endif
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t hint_ns = CSF_WAIT_HINT_NS) {
    // try_wait with an explicit suspend-time hint: the thread is parked by the hardware for up to
    // hint_ns per attempt instead of re-issuing the poll
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITB_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONEB_%=;\n\t"
        "bra WAITB_%=;\n\t"
        "DONEB_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
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
    k.sg2s = (T)(fp->sigma_2 / kappa / 1.4142135623730951);
    k.sg3s = (T)(fp->sigma_3 / kappa / 1.4142135623730951);
    k.e0 = (T)fp->e_0;
    k.e1 = (T)fp->e_1;
    k.qa = (T)(1.0 - fp->e_0 * fp->e_0);
    k.qb = (T)(2.0 * fp->e_0 * fp->e_1);
    k.qc = (T)(-fp->e_1 * fp->e_1);
    k.ncosH = (fp->hfov * 0.5 >= CSF_PI) ? (T)2 : (T)(-cos(fp->hfov * 0.5));
    k.tiny = (T)(is_f32 ? 1e-6 : 1e-200);
    if (fp->field_kind == 1) k.sg0 = (T)(kappa / fp->p_decay);   // v0.1 Bicycle field: exponent per payload unit
    return k;
}
// amplitude the reduced sums are scaled with
inline double field_amplitude(const CsfFieldParams* fp) { return fp->field_kind == 1 ? fp->p_0 / fp->p_decay : fp->f_0; }


}  // namespace
