// K1 (tiled): all-pairs repulsive force with hierarchical culling.
//
// The masked sum  Frep_j = sum_i mask(i,j) F(i -> j)  (reference intersection.py:788-843,
// vehicle.py:1560-1648) has two sources of exact or negligible zeros:
//   * ~2/3 of the ordered pairs are masked by the *target's* field of view
//     (intersection.py:733-736; hfov = 2 pi / 3);
//   * |F(i -> j)| = f_0 exp(-rho q / sigma) (vehicle.py:1613-1648): beyond rho = d_cut every
//     contribution is below 2^-cutoff_log2 f_0 (f32 build only; d_cut ~ 160 m with the default
//     parameters; the f64 verification build never truncates).
// Both are turned into skipped work at three levels:
//   1. sources live in a spatially sorted copy (Morton order, re-sorted by the host every few
//      steps) cut into tiles of 64 and chunks of 16 tiles, each with a bounding circle; targets are
//      visited in Morton order too, in blocks of 8 warps x tpw targets with a bounding circle;
//   2. the producer warp of a CTA streams only the chunks that can be within d_cut of the
//      target block (lane-parallel circle/circle test) into a shared-memory ring with TMA bulk
//      copies + mbarriers;
//   3. a consumer warp works on ONE target at a time: its lanes test one tile each against the
//      target's view cone expanded by the tile radius and the cut-off distance (warp-uniform
//      decision after a ballot), then evaluate the surviving tiles two sources per lane with the
//      same pair_eval as the dense kernel (which still applies the exact per-pair mask).
// Culled pairs are pairs whose contribution is exactly 0 (mask) or < 2^-cutoff_log2 f_0 (f32), so
// the result equals the dense kernel's up to the order of summation and that bound.
// Work items (target block x chunk group) are handed out dynamically (atomic counter); partial
// sums per chunk group are reduced in fixed order (deterministic, no float atomics).
#include "csf_common.cuh"
#include "csf_pair_common.cuh"
#include <stdlib.h>

namespace {

#ifndef CSF_TILED_WARPS
#define CSF_TILED_WARPS 8        // consumer warps per CTA
#endif
#ifndef CSF_TILED_MINB
#define CSF_TILED_MINB 3
#endif
constexpr int kTW = CSF_TILED_WARPS;
constexpr int kTThreads = (kTW + 1) * 32;
constexpr int kMaxTPW = 16;              // targets per warp per item (runtime tpw <= kMaxTPW)
constexpr int kTB = kTW * kMaxTPW;       // targets per block, at most
constexpr int kTileS = 64;               // sources per tile (2 per lane)
constexpr int kCT = 16;                  // tiles per chunk = one shared-memory stage
constexpr int kCS = kCT * kTileS;        // sources per chunk
constexpr int kTMaxGroups = 64;
template <typename T> struct Stages { static constexpr int n = 3; };

template <typename T> struct Tile;
template <> struct __align__(16) Tile<float> { int32_t cx, cy; float R; int32_t cnt; };
template <> struct __align__(16) Tile<double> { double cx, cy, R; int64_t cnt; };

// Sorted-copy layout, per tile of 64 (lane l owns sources l and l + 32 of the tile):
//   SrcA[32] = {x_l, x_{l+32}, y_l, y_{l+32}}   then   SrcB[32] = {c_l, c_{l+32}, s_l, s_{l+32}}
// so a lane fetches both of its sources with two 16-byte (f64: 32-byte) conflict-free loads.
template <typename T> struct SrcA;
template <> struct __align__(16) SrcA<float> { int32_t x0, x1, y0, y1; };
template <> struct __align__(16) SrcA<double> { double x0, x1, y0, y1; };
template <typename T> struct __align__(16) SrcB { T c0, c1, s0, s1; };
template <typename T> struct TileBytes { static constexpr size_t v = (size_t)kTileS * sizeof(Xycs<T>); };

constexpr int kLobeBins = 64;
template <typename T> struct CullConst {
    T ca, sa;    // cos / sin of hfov/2
    T dmax;      // cut-off distance d_cut in payload units (huge: never)
    // reach of a source's field in the direction phi (measured from its heading), as a step function of
    // cos(phi): lobe[b] >= max distance at which |F| >= 2^-cutoff_log2 f_0 for any phi' with
    // cos(phi') <= -1 + (b+1) 2/kLobeBins and any heading difference (payload units; huge: never)
    T lobe[kLobeBins];
};

__device__ __forceinline__ void split(const Xycs<float>& a, const Xycs<float>& b, SrcA<float>& A, SrcB<float>& B) {
    A.x0 = a.xq; A.x1 = b.xq; A.y0 = a.yq; A.y1 = b.yq;
    B.c0 = a.c; B.c1 = b.c; B.s0 = a.s; B.s1 = b.s;
}
__device__ __forceinline__ void split(const Xycs<double>& a, const Xycs<double>& b, SrcA<double>& A, SrcB<double>& B) {
    A.x0 = a.x; A.x1 = b.x; A.y0 = a.y; A.y1 = b.y;
    B.c0 = a.c; B.c1 = b.c; B.s0 = a.s; B.s1 = b.s;
}
__device__ __forceinline__ Xycs<float> src0(const SrcA<float>& A, const SrcB<float>& B) { return {A.x0, A.y0, B.c0, B.s0}; }
__device__ __forceinline__ Xycs<float> src1(const SrcA<float>& A, const SrcB<float>& B) { return {A.x1, A.y1, B.c1, B.s1}; }
__device__ __forceinline__ Xycs<double> src0(const SrcA<double>& A, const SrcB<double>& B) { return {A.x0, A.y0, B.c0, B.s0}; }
__device__ __forceinline__ Xycs<double> src1(const SrcA<double>& A, const SrcB<double>& B) { return {A.x1, A.y1, B.c1, B.s1}; }

// Both sources of this lane in one tile against target tg: (ax, ay) and (bx, by) each take one source.
// f32: packed FFMA2 evaluation, the two halves of (ax, ay) are the two sources' partial sums.
template <bool P2R> struct TileAcc32 {
    F2 x, y;
    __device__ __forceinline__ TileAcc32() : x(splat(0.f)), y(splat(0.f)) {}
    __device__ __forceinline__ void eval(const SrcA<float>& A, const SrcB<float>& B, const Tgt<float>& tg,
                                         const PairConst<float>& k) {
        pair_eval2<P2R>(A.x0, A.x1, A.y0, A.y1, B.c0, B.c1, B.s0, B.s1, tg, k, x, y);
    }
    __device__ __forceinline__ void merge(const TileAcc32& o) { x = add2(x, o.x); y = add2(y, o.y); }
    __device__ __forceinline__ void total(float& sx, float& sy) const {
        float a, b;
        up(x, a, b); sx = a + b;
        up(y, a, b); sy = a + b;
    }
};
template <bool P2R> struct TileAcc64 {
    double x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    __device__ __forceinline__ void eval(const SrcA<double>& A, const SrcB<double>& B, const Tgt<double>& tg,
                                         const PairConst<double>& k) {
        pair_eval<double, P2R>(src0(A, B), tg, k, x0, y0);
        pair_eval<double, P2R>(src1(A, B), tg, k, x1, y1);
    }
    __device__ __forceinline__ void merge(const TileAcc64& o) { x0 += o.x0; y0 += o.y0; x1 += o.x1; y1 += o.y1; }
    __device__ __forceinline__ void total(double& sx, double& sy) const { sx = x0 + x1; sy = y0 + y1; }
};
template <typename T, bool P2R> struct TileAccSel;
template <bool P2R> struct TileAccSel<float, P2R> { typedef TileAcc32<P2R> type; };
template <bool P2R> struct TileAccSel<double, P2R> { typedef TileAcc64<P2R> type; };

// ---- bounding circles ---------------------------------------------------------------------------
__device__ __forceinline__ void pad_entry(Xycs<float>& e) { e.xq = 1 << 30; e.yq = 1 << 30; e.c = 1.f; e.s = 0.f; }
__device__ __forceinline__ void pad_entry(Xycs<double>& e) { e.x = 1e150; e.y = 1e150; e.c = 1.0; e.s = 0.0; }

// running bounding box of payload positions (integer for the Q-format payload: exact)
template <typename T> struct BBox;
template <> struct BBox<float> {
    int xmin = INT32_MAX, xmax = INT32_MIN, ymin = INT32_MAX, ymax = INT32_MIN;
    __device__ __forceinline__ void add(int x, int y) { xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y); }
    __device__ __forceinline__ void add(const Xycs<float>& e) { add(e.xq, e.yq); }
    __device__ __forceinline__ void warp_reduce() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
    }
    __device__ __forceinline__ Tile<float> circle(int64_t cnt) const {
        Tile<float> t;
        if (cnt <= 0) { t.cx = 1 << 30; t.cy = 1 << 30; t.R = 0.f; t.cnt = 0; return t; }
        t.cx = (int)(((int64_t)xmin + xmax) >> 1);
        t.cy = (int)(((int64_t)ymin + ymax) >> 1);
        const float hx = (float)((int64_t)xmax - xmin) * 0.5f + 1.f, hy = (float)((int64_t)ymax - ymin) * 0.5f + 1.f;
        t.R = sqrtf(hx * hx + hy * hy) * 1.000001f + 1.f;
        t.cnt = (int32_t)cnt;
        return t;
    }
};
template <> struct BBox<double> {
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    __device__ __forceinline__ void add(double x, double y) { xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y); }
    __device__ __forceinline__ void add(const Xycs<double>& e) { add(e.x, e.y); }
    __device__ __forceinline__ void warp_reduce() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
    }
    __device__ __forceinline__ Tile<double> circle(int64_t cnt) const {
        Tile<double> t;
        if (cnt <= 0) { t.cx = 1e150; t.cy = 1e150; t.R = 0.0; t.cnt = 0; return t; }
        t.cx = 0.5 * (xmin + xmax);
        t.cy = 0.5 * (ymin + ymax);
        const double hx = 0.5 * (xmax - xmin), hy = 0.5 * (ymax - ymin);
        t.R = sqrt(hx * hx + hy * hy) * (1.0 + 1e-12) + 1e-9;
        t.cnt = cnt;
        return t;
    }
};

// One CTA per chunk, one warp per tile: gather the tile's 64 sources through `perm` (nullptr:
// identity), write them in the SrcA/SrcB layout with the tile's bounding circle, then combine the 16
// tile boxes into the chunk's bounding circle.  Entries past n are padded with a far-away sentinel
// that contributes exactly 0 and is not part of any bounding circle.
template <typename T>
__global__ void __launch_bounds__(kCT * 32)
tile_sources_kernel(const Xycs<T>* __restrict__ xycs, int64_t n, const int64_t* __restrict__ perm,
                    unsigned char* __restrict__ sorted, Tile<T>* __restrict__ tiles, int64_t n_tiles) {
    __shared__ __align__(16) unsigned char boxes_raw[kCT * sizeof(BBox<T>)];
    BBox<T>* boxes = reinterpret_cast<BBox<T>*>(boxes_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t c = blockIdx.x, t = c * kCT + w;
    BBox<T> bb;
    if (t < n_tiles) {
        const int64_t i0 = t * kTileS + lane, i1 = i0 + 32;
        Xycs<T> a, b;
        const bool va = i0 < n, vb = i1 < n;
        if (va) a = xycs[perm ? perm[i0] : i0]; else pad_entry(a);
        if (vb) b = xycs[perm ? perm[i1] : i1]; else pad_entry(b);
        SrcA<T> A;
        SrcB<T> B;
        split(a, b, A, B);
        unsigned char* base = sorted + (size_t)t * TileBytes<T>::v;
        reinterpret_cast<SrcA<T>*>(base)[lane] = A;
        reinterpret_cast<SrcB<T>*>(base + 32 * sizeof(SrcA<T>))[lane] = B;
        if (va) bb.add(a);
        if (vb) bb.add(b);
        bb.warp_reduce();
        const int64_t rem = n - t * kTileS;
        if (lane == 0) tiles[t] = bb.circle(rem < kTileS ? rem : kTileS);
    }
    if (lane == 0) boxes[w] = bb;
    __syncthreads();
    if (w == 0) {
        BBox<T> cb;
        if (lane < kCT) cb = boxes[lane];
        cb.warp_reduce();
        const int64_t cnt = min(n, (c + 1) * (int64_t)kCS) - c * (int64_t)kCS;
        if (lane == 0) tiles[n_tiles + c] = cb.circle(cnt);
    }
}

// One warp per target block: bounding circle of targets tgt[perm[b*group .. (b+1)*group)).
template <typename T>
__global__ void block_bounds_kernel(const Xycs<T>* __restrict__ tgt, const int64_t* __restrict__ perm, int64_t n,
                                    int group, Tile<T>* __restrict__ blocks, int64_t n_blocks,
                                    unsigned int* __restrict__ item_counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *item_counter = 0;   // the pair kernel's dynamic item counter
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= n_blocks) return;
    BBox<T> bb;
    const int64_t i_end = min(n, (b + 1) * (int64_t)group);
    for (int64_t i = b * (int64_t)group + lane; i < i_end; i += 32) bb.add(tgt[perm ? perm[i] : i]);
    bb.warp_reduce();
    if (lane == 0) blocks[b] = bb.circle(i_end - b * (int64_t)group);
}

// Spatial sort key of a payload position: index along a 2^16 x 2^16 Hilbert curve (the host sorts the
// keys; any order gives correct results).  Consecutive runs of a Hilbert order are compact -- the mean
// bounding radius of a 64-source tile is 27 m at 4 m spacing against 40 m for a Morton order, which
// is 23 % fewer evaluated pairs after culling.
// (x0, y0) = lower corner of the crowd's bounding box, inv_cell = 65535 / its larger side, both in
// payload units: the curve then covers exactly the occupied region.  With a fixed, larger key domain
// the few road users that drift across one of its quadrant boundaries sort far away from their
// neighbours and blow up the bounding circles of the tiles and target blocks they land in.
__device__ __forceinline__ void key_xy(const Xycs<float>& e, double x0, double y0, double inv_cell, uint32_t& kx,
                                       uint32_t& ky) {
    kx = (uint32_t)fmin(fmax(((double)e.xq - x0) * inv_cell, 0.0), 65535.0);
    ky = (uint32_t)fmin(fmax(((double)e.yq - y0) * inv_cell, 0.0), 65535.0);
}
__device__ __forceinline__ void key_xy(const Xycs<double>& e, double x0, double y0, double inv_cell, uint32_t& kx,
                                       uint32_t& ky) {
    kx = (uint32_t)fmin(fmax((e.x - x0) * inv_cell, 0.0), 65535.0);
    ky = (uint32_t)fmin(fmax((e.y - y0) * inv_cell, 0.0), 65535.0);
}
__device__ __forceinline__ uint32_t hilbert_index(uint32_t x, uint32_t y) {
    uint32_t d = 0;
#pragma unroll
    for (uint32_t s = 1u << 15; s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = 65535u - x; y = 65535u - y; }
            const uint32_t t = x; x = y; y = t;
        }
    }
    return d;
}
template <typename T>
__global__ void morton_kernel(const Xycs<T>* __restrict__ xycs, int64_t n, double x0, double y0, double inv_cell,
                              const double* __restrict__ box, int64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (box) {      // {xmin, xmax, ymin, ymax} on the device: no host round trip
        x0 = box[0];
        y0 = box[2];
        inv_cell = 65535.0 / fmax(fmax(box[1] - box[0], box[3] - box[2]), 1e-300);
    }
    uint32_t kx, ky;
    key_xy(xycs[i], x0, y0, inv_cell, kx, ky);
    keys[i] = (int64_t)hilbert_index(kx, ky);
}

// ---- cull tests ----------------------------------------------------------------------------------
// Tile circle (c, R) against target j's view cone: u = c - p_j in the target's frame,
// along = u . h_j, cross = u x h_j.  Every point of the circle is outside the (closed) cone of half
// angle a if  |cross| cos a - along sin a > R  (distance to the supporting half-plane; also valid
// for a >= 90 deg, where it states that the circle lies in the complementary cone).  p2r additionally
// hides sources to the left of the heading.  Beyond d_cut + R every source of the tile is negligible.
__device__ __forceinline__ void tile_delta(const Tile<float>& t, const Tgt<float>& g, float& dx, float& dy) {
    dx = (float)(t.cx - g.xq);
    dy = (float)(t.cy - g.yq);
}
__device__ __forceinline__ void tile_delta(const Tile<double>& t, const Tgt<double>& g, double& dx, double& dy) {
    dx = t.cx - g.x;
    dy = t.cy - g.y;
}
template <typename T, bool P2R>
__device__ __forceinline__ bool tile_visible(const Tile<T>& t, const Tgt<T>& g, const CullConst<T>& cc) {
    T dx, dy;
    tile_delta(t, g, dx, dy);
    const T along = fma(dy, g.s, dx * g.c);
    const T cross = fma(dy, g.c, -(dx * g.s));   // > 0: tile centre to the left of the heading
    const T R = (T)t.R;
    bool vis = fma(fabs(cross), cc.ca, -(along * cc.sa)) <= R;
    if (P2R) vis = vis && (cross <= R);
    const T far = R + cc.dmax;
    vis = vis && (fma(dx, dx, dy * dy) <= far * far);
    return vis;
}
// chunk circle against target-block circle: can any pair be within d_cut?
__device__ __forceinline__ bool circles_near(const Tile<float>& a, const Tile<float>& b, float dmax) {
    const float dx = (float)((int64_t)a.cx - b.cx), dy = (float)((int64_t)a.cy - b.cy);
    const float far = (a.R + b.R + dmax) * 1.000001f;
    return a.cnt > 0 && b.cnt > 0 && fmaf(dx, dx, dy * dy) <= far * far;
}
__device__ __forceinline__ bool circles_near(const Tile<double>& a, const Tile<double>& b, double dmax) {
    const double dx = a.cx - b.cx, dy = a.cy - b.cy;
    const double far = a.R + b.R + dmax;
    return a.cnt > 0 && b.cnt > 0 && fma(dx, dx, dy * dy) <= far * far;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int32_t pos_x(const Xycs<float>& e) { return e.xq; }
__device__ __forceinline__ int32_t pos_y(const Xycs<float>& e) { return e.yq; }
__device__ __forceinline__ double pos_x(const Xycs<double>& e) { return e.x; }
__device__ __forceinline__ double pos_y(const Xycs<double>& e) { return e.y; }

// ---- lobe filter: can source (x, y, heading c, s) matter for ANY target inside the block circle? ----
// d = distance source -> block centre, phi0 = its direction seen from the source's heading, delta =
// half angle under which the source sees the circle.  Every point of the circle lies in a direction
// |phi| >= |phi0| - delta at a distance >= d - R_b, and the reach is tabulated as a non-decreasing step
// function of cos(phi): the source cannot matter if d - R_b exceeds the reach at cos(|phi0| - delta).
__device__ __forceinline__ void block_delta(const Tile<float>& b, int32_t x, int32_t y, float& dx, float& dy) {
    dx = (float)(b.cx - x);
    dy = (float)(b.cy - y);
}
__device__ __forceinline__ void block_delta(const Tile<double>& b, double x, double y, double& dx, double& dy) {
    dx = b.cx - x;
    dy = b.cy - y;
}
template <typename T, typename P>
__device__ __forceinline__ bool lobe_reaches(const Tile<T>& blk, P x, P y, T c, T s, const T* __restrict__ lobe, T tiny) {
    T dx, dy;
    block_delta(blk, x, y, dx, dy);
    const T d2 = fma(dy, dy, fma(dx, dx, tiny));
    const T rinv = M<T>::rsqrt(d2);
    const T d = d2 * rinv;
    const T Rb = (T)blk.R;
    const T cphi = fma(dy, s, dx * c) * rinv;
    const T sphi = fabs(fma(dy, c, -(dx * s))) * rinv;
    const T sdel = fmin(Rb * rinv, (T)1);
    const T cdel = M<T>::sqrt(fmax(fma(-sdel, sdel, (T)1), (T)0));
    T cmin = fma(cphi, cdel, sphi * sdel);                 // cos(|phi0| - delta)
    cmin = (cphi >= cdel) ? (T)1 : cmin;                   // the circle straddles the heading direction
    const int bin = min(max((int)((cmin + (T)1.00002) * (T)(kLobeBins / 2)), 0), kLobeBins - 1);
    return d - Rb <= lobe[bin] * (T)1.0001;
}

// both sources of a lane at once; f32: the FP32 arithmetic packed two-wide (FFMA2), same operations
__device__ __forceinline__ void lobe_reaches2(const Tile<float>& blk, const SrcA<float>& A, const SrcB<float>& B,
                                              const float* __restrict__ lobe, float tiny, bool& r0, bool& r1) {
    const F2 DX = pk((float)(blk.cx - A.x0), (float)(blk.cx - A.x1));
    const F2 DY = pk((float)(blk.cy - A.y0), (float)(blk.cy - A.y1));
    const F2 D2 = fma2(DY, DY, fma2(DX, DX, splat(tiny)));
    float a0, a1;
    up(D2, a0, a1);
    const F2 RINV = pk(M<float>::rsqrt(a0), M<float>::rsqrt(a1));
    const F2 D = mul2(D2, RINV);
    const F2 SC = pk(B.c0, B.c1), SS = pk(B.s0, B.s1);
    const F2 CPHI = mul2(fma2(DY, SS, mul2(DX, SC)), RINV);
    const F2 SPHI = mul2(abs2(fma2(DY, SC, neg2(mul2(DX, SS)))), RINV);
    const F2 SDELr = mul2(splat(blk.R), RINV);
    up(SDELr, a0, a1);
    const F2 SDEL = pk(fminf(a0, 1.f), fminf(a1, 1.f));
    const F2 CD2 = fma2(neg2(SDEL), SDEL, splat(1.f));
    up(CD2, a0, a1);
    const float cd0 = M<float>::sqrt(fmaxf(a0, 0.f)), cd1 = M<float>::sqrt(fmaxf(a1, 0.f));
    const F2 CMIN = fma2(CPHI, pk(cd0, cd1), mul2(SPHI, SDEL));
    float c0, c1, p0, p1, d0, d1;
    up(CMIN, c0, c1);
    up(CPHI, p0, p1);
    up(D, d0, d1);
    c0 = (p0 >= cd0) ? 1.f : c0;
    c1 = (p1 >= cd1) ? 1.f : c1;
    const int b0 = min(max((int)((c0 + 1.00002f) * (float)(kLobeBins / 2)), 0), kLobeBins - 1);
    const int b1 = min(max((int)((c1 + 1.00002f) * (float)(kLobeBins / 2)), 0), kLobeBins - 1);
    r0 = d0 - blk.R <= lobe[b0] * 1.0001f;
    r1 = d1 - blk.R <= lobe[b1] * 1.0001f;
}
__device__ __forceinline__ void lobe_reaches2(const Tile<double>& blk, const SrcA<double>& A, const SrcB<double>& B,
                                              const double* __restrict__ lobe, double tiny, bool& r0, bool& r1) {
    r0 = lobe_reaches<double, double>(blk, A.x0, A.y0, B.c0, B.s0, lobe, tiny);
    r1 = lobe_reaches<double, double>(blk, A.x1, A.y1, B.c1, B.s1, lobe, tiny);
}

// named barrier 1 among the consumer warps only (the producer warp never joins)
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTW * 32) : "memory"); }

// ---- the tiled pair kernel ---------------------------------------------------------------------------
// item -> (target block tb = item % n_tblocks, chunk group cg = item / n_tblocks); a group is
// `group_chunks` consecutive chunks.  partial[cg][target][2].
// Stage protocol: the producer fills hdr[stage] = {item, chunk} and the stage's bytes; chunk == -1
// closes the item (consumers write their partial sums), chunk == -2 ends the kernel.
template <typename T, bool P2R>
__global__ void __launch_bounds__(kTThreads, sizeof(T) == 4 ? CSF_TILED_MINB : 1)
pair_tiled_kernel(const unsigned char* __restrict__ sorted, const Tile<T>* __restrict__ tiles, int64_t n_tiles,
                  const Xycs<T>* __restrict__ tgt, const int64_t* __restrict__ tgt_perm, int64_t n_tgt,
                  const Tile<T>* __restrict__ tblocks, PairConst<T> k, CullConst<T> cc, T* __restrict__ partial,
                  int tpw, int group_chunks, int n_groups, int n_tblocks, unsigned int* __restrict__ counter,
                  unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kStages = Stages<T>::n;
    constexpr size_t kStageBytes = (size_t)kCS * sizeof(Xycs<T>) + (size_t)kCT * sizeof(Tile<T>);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    int4* hdr = reinterpret_cast<int4*>(empty + kStages);                           // {item, chunk, -, -}
    unsigned char* sbuf = reinterpret_cast<unsigned char*>(hdr + kStages);          // survivors: one chunk in tile layout
    Tile<T>* dtile = reinterpret_cast<Tile<T>*>(sbuf + (size_t)kCS * sizeof(Xycs<T>));      // [kCT] their circles
    Xycs<T>* btgt = reinterpret_cast<Xycs<T>*>(dtile + kCT);                        // [kTB] targets of the block
    T* bacc = reinterpret_cast<T*>(btgt + kTB);                                     // [kTB][2] their sums
    T* lobe = bacc + kTB * 2;                                                       // [kLobeBins] reach table
    uint32_t* wlist = reinterpret_cast<uint32_t*>(lobe + kLobeBins);                // [kTB] (q << 16) | tile mask
    uint2* fmask = reinterpret_cast<uint2*>(wlist + kTB);                           // [kStages][kCT] filter ballots
    int* lctl = reinterpret_cast<int*>(fmask + kStages * kCT);                      // list length, next entry

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kLobeBins; i += blockDim.x) lobe[i] = cc.lobe[i];
    __syncthreads();

    const unsigned int n_items = (unsigned int)n_tblocks * (unsigned int)n_groups;
    const int64_t n_chunks = (n_tiles + kCT - 1) / kCT;
    const Tile<T>* chunks = tiles + n_tiles;

    if (warp == kTW) {
        // ===== producer warp: fetch items, cull chunks against the target block, stream the rest =====
        uint32_t it = 0;
        for (;;) {
            unsigned int item = 0;
            if (lane == 0) item = atomicAdd(counter, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= n_items) break;
            const int tb = (int)(item % (unsigned int)n_tblocks), cg = (int)(item / (unsigned int)n_tblocks);
            const Tile<T> tbr = tblocks[tb];
            const int64_t c_begin = (int64_t)cg * group_chunks, c_end = min(n_chunks, c_begin + group_chunks);
            for (int64_t c0 = c_begin; c0 < c_end; c0 += 32) {
                const int64_t c = c0 + lane;
                bool near = false;
                if (c < c_end) near = circles_near(chunks[c], tbr, cc.dmax);
                uint32_t m = __ballot_sync(0xffffffffu, near);
                while (m) {
                    const int64_t ch = c0 + (__ffs(m) - 1);
                    m &= m - 1;
                    const int stage = it % kStages;
                    const uint32_t phase = (it / kStages) & 1;
                    if (lane == 0) {
                        mbar_wait_backoff(&empty[stage], phase ^ 1);
                        const int64_t t0 = ch * kCT;
                        const uint32_t nt = (uint32_t)min((int64_t)kCT, n_tiles - t0);
                        const uint32_t bsrc = nt * (uint32_t)TileBytes<T>::v, btile = nt * (uint32_t)sizeof(Tile<T>);
                        unsigned char* base = smem_raw + stage * kStageBytes;
                        hdr[stage] = make_int4((int)item, (int)ch, 0, 0);
                        mbar_expect_tx(&full[stage], bsrc + btile);
                        tma_bulk_g2s(base, sorted + (size_t)t0 * TileBytes<T>::v, bsrc, &full[stage]);
                        tma_bulk_g2s(base + (size_t)kCS * sizeof(Xycs<T>), tiles + t0, btile, &full[stage]);
                    }
                    ++it;
                }
            }
            {   // close the item
                const int stage = it % kStages;
                const uint32_t phase = (it / kStages) & 1;
                if (lane == 0) {
                    mbar_wait_backoff(&empty[stage], phase ^ 1);
                    hdr[stage] = make_int4((int)item, -1, 0, 0);
                    mbar_arrive(&full[stage]);
                }
                ++it;
            }
            __syncwarp();
        }
        if (lane == 0) {
            const int stage = it % kStages;
            const uint32_t phase = (it / kStages) & 1;
            mbar_wait_backoff(&empty[stage], phase ^ 1);
            hdr[stage] = make_int4(0, -2, 0, 0);
            mbar_arrive(&full[stage]);
        }
        return;
    }

    // ===== consumer warps =====
    // Per streamed chunk: (1) filter -- every source is tested against the target block's circle with
    // the lobe test (lobe_reaches): most sources of a chunk near the block cannot matter for any of its
    // targets, because a source's field reaches far only in a narrow range of directions; the
    // survivors are appended, in stream order, to a shared-memory buffer that holds one chunk's worth
    // (16 dynamic tiles of 64) in the same tile layout.  When the buffer would overflow, and when the
    // item closes, it is flushed: (2) bounding circles of the dynamic tiles, (3) cull -- every warp
    // tests its own targets' view cones against the 16 circles, two targets per pass, and appends
    // (target, tile mask) entries to a work list, (4) evaluate -- warps take entries from the list one
    // at a time (shared counter; the targets' headings make the per-target work very uneven).
    // One warp per (target, flush), flushes separated by barriers: deterministic sums, no float atomics.
    uint32_t it = 0;
    unsigned long long n_eval = 0;
    int cur = -1, cg = 0, nq = 0, count = 0;
    const int q0 = warp * tpw;           // this warp's targets in the block: [q0, q0 + nq)
    long long myj = -1;
    Tile<T> blk;                         // bounding circle of the block's targets
    typedef decltype(SrcA<T>().x0) P;    // payload position type

    auto write_entry = [&](int kidx, P x, P y, T c, T s) {
        unsigned char* tb = sbuf + (size_t)(kidx >> 6) * TileBytes<T>::v;
        const int r = kidx & 63, l = r & 31, h = r >> 5;
        P* pa = reinterpret_cast<P*>(reinterpret_cast<SrcA<T>*>(tb) + l);
        T* pb = reinterpret_cast<T*>(reinterpret_cast<SrcB<T>*>(tb + 32 * sizeof(SrcA<T>)) + l);
        pa[h] = x;
        pa[2 + h] = y;
        pb[h] = c;
        pb[2 + h] = s;
    };
    auto cull = [&](int nt) {
        const int tl = lane & (kCT - 1), half = lane >> 4;
        Tile<T> mytile;
        if (tl < nt) mytile = dtile[tl];
        for (int i0 = 0; i0 < nq; i0 += 2) {          // warp-uniform trip count (ballot inside)
            const int i = i0 + half;
            const bool valid = (i < nq) && (tl < nt);
            const Xycs<T> te = btgt[q0 + (i < nq ? i : 0)];
            const Tgt<T> tg = *reinterpret_cast<const Tgt<T>*>(&te);
            const bool v = valid && tile_visible<T, P2R>(mytile, tg, cc);
            const uint32_t m2 = __ballot_sync(0xffffffffu, v);
            const uint32_t m = half ? (m2 >> 16) : (m2 & 0xffffu);
            if (tl == 0 && m) {
                wlist[atomicAdd(&lctl[0], 1)] = ((uint32_t)(q0 + i) << 16) | m;
                if (stats) n_eval += (unsigned long long)__popc(m) * kTileS;
            }
        }
    };
    auto evaluate = [&]() {
        const unsigned char* base = sbuf;
        const int n_list = lctl[0];
        for (;;) {
            int e = 0;
            if (lane == 0) e = atomicAdd(&lctl[1], 1);
            e = __shfl_sync(0xffffffffu, e, 0);
            if (e >= n_list) break;
            const uint32_t ent = wlist[e];
            const int q = (int)(ent >> 16);
            uint32_t mask = ent & 0xffffu;
            const Xycs<T> te = btgt[q];
            const Tgt<T> tg = *reinterpret_cast<const Tgt<T>*>(&te);
            typename TileAccSel<T, P2R>::type acc0, acc1;
            // two surviving tiles per iteration: four independent pair evaluations in flight
            while (mask & (mask - 1)) {
                const int t0 = __ffs(mask) - 1;
                mask &= mask - 1;
                const int t1 = __ffs(mask) - 1;
                mask &= mask - 1;
                const unsigned char* p0 = base + (size_t)t0 * TileBytes<T>::v;
                const unsigned char* p1 = base + (size_t)t1 * TileBytes<T>::v;
                const SrcA<T> A0 = reinterpret_cast<const SrcA<T>*>(p0)[lane];
                const SrcB<T> B0 = reinterpret_cast<const SrcB<T>*>(p0 + 32 * sizeof(SrcA<T>))[lane];
                const SrcA<T> A1 = reinterpret_cast<const SrcA<T>*>(p1)[lane];
                const SrcB<T> B1 = reinterpret_cast<const SrcB<T>*>(p1 + 32 * sizeof(SrcA<T>))[lane];
                acc0.eval(A0, B0, tg, k);
                acc1.eval(A1, B1, tg, k);
            }
            if (mask) {
                const int t0 = __ffs(mask) - 1;
                const unsigned char* p0 = base + (size_t)t0 * TileBytes<T>::v;
                const SrcA<T> A0 = reinterpret_cast<const SrcA<T>*>(p0)[lane];
                const SrcB<T> B0 = reinterpret_cast<const SrcB<T>*>(p0 + 32 * sizeof(SrcA<T>))[lane];
                acc0.eval(A0, B0, tg, k);
            }
            acc0.merge(acc1);
            T ax, ay;
            acc0.total(ax, ay);
            // both sums in one butterfly: lanes 0-15 reduce x, lanes 16-31 reduce y
            const bool upper = lane >= 16;
            T mine = upper ? ay : ax;
            const T other = upper ? ax : ay;
            mine += __shfl_xor_sync(0xffffffffu, other, 16);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            if ((lane & 15) == 0) bacc[q * 2 + (lane >> 4)] += mine;
        }
    };
    // evaluate the `count` survivors in the buffer against every target of the block
    auto flush = [&]() {
        consumer_barrier();                                   // every append has landed
        const int n_dt = (count + kTileS - 1) / kTileS;
        {   // pad the last dynamic tile with far-away sentinels (they contribute exactly 0)
            Xycs<T> pad;
            pad_entry(pad);
            for (int kidx = count + (int)threadIdx.x; kidx < n_dt * kTileS; kidx += kTW * 32)
                write_entry(kidx, pos_x(pad), pos_y(pad), pad.c, pad.s);
        }
        if (threadIdx.x == 0) { lctl[0] = 0; lctl[1] = 0; }
        for (int t = warp; t < n_dt; t += kTW) {              // circles of the dynamic tiles
            const int valid = min(kTileS, count - t * kTileS);
            const SrcA<T> A = reinterpret_cast<const SrcA<T>*>(sbuf + (size_t)t * TileBytes<T>::v)[lane];
            BBox<T> bb;
            if (lane < valid) bb.add(A.x0, A.y0);
            if (lane + 32 < valid) bb.add(A.x1, A.y1);
            bb.warp_reduce();
            if (lane == 0) dtile[t] = bb.circle(valid);
        }
        consumer_barrier();
        cull(n_dt);
        consumer_barrier();
        evaluate();
        consumer_barrier();                                   // the buffer may be overwritten again
    };

    for (;;) {
        const int stage = it % kStages;
        mbar_wait(&full[stage], (it / kStages) & 1);
        const int item = hdr[stage].x, chunk = hdr[stage].y;
        if (chunk == -2) break;
        if (item != cur) {
            // open the item: lane i < nq of each warp fetches the warp's target i
            cur = item;
            const int tb = cur % n_tblocks;
            cg = cur / n_tblocks;
            blk = tblocks[tb];
            const int64_t t_first = ((int64_t)tb * kTW + warp) * tpw;
            const int64_t left = n_tgt - t_first;
            nq = (int)(left < 0 ? 0 : (left < tpw ? left : tpw));
            myj = -1;
            count = 0;
            if (lane < nq) {
                myj = tgt_perm ? tgt_perm[t_first + lane] : (t_first + lane);
                btgt[q0 + lane] = tgt[myj];
            }
            if (lane < 2 * nq) bacc[q0 * 2 + lane] = (T)0;
            __syncwarp();
        }
        if (chunk == -1) {
            // close the item: evaluate what is left in the buffer, then write this warp's targets
            if (count > 0) flush();
            count = 0;
            const long long jj = __shfl_sync(0xffffffffu, myj, (lane >> 1) & (kMaxTPW - 1));
            if ((lane >> 1) < nq) partial[((size_t)cg * n_tgt + (size_t)jj) * 2 + (lane & 1)] = bacc[q0 * 2 + lane];
        } else {
            const int nt = (int)min((int64_t)kCT, n_tiles - (int64_t)chunk * kCT);
            const unsigned char* base = smem_raw + stage * kStageBytes;
            const Tile<T>* trec = reinterpret_cast<const Tile<T>*>(base + (size_t)kCS * sizeof(Xycs<T>));
            uint2* fm = fmask + stage * kCT;
            // (1a) filter: this warp's tiles of the chunk; tiles wholly out of reach of the block first
            uint32_t tnear;
            {
                const int tl = lane & (kCT - 1);
                tnear = __ballot_sync(0xffffffffu, tl < nt && circles_near(trec[tl < nt ? tl : 0], blk, cc.dmax));
            }
            for (int t = warp; t < kCT; t += kTW) {
                uint32_t b0 = 0, b1 = 0;
                if ((tnear >> t) & 1u) {
                    const int valid = (int)trec[t].cnt;
                    const SrcA<T> A = reinterpret_cast<const SrcA<T>*>(base + (size_t)t * TileBytes<T>::v)[lane];
                    const SrcB<T> B = reinterpret_cast<const SrcB<T>*>(base + (size_t)t * TileBytes<T>::v + 32 * sizeof(SrcA<T>))[lane];
                    bool p0, p1;
                    lobe_reaches2(blk, A, B, lobe, k.tiny, p0, p1);
                    b0 = __ballot_sync(0xffffffffu, p0 && (lane < valid));
                    b1 = __ballot_sync(0xffffffffu, p1 && (lane + 32 < valid));
                }
                if (lane == 0) fm[t] = make_uint2(b0, b1);
            }
            consumer_barrier();
            // (1b) append in stream order: prefix over the chunk's 16 tiles (every warp computes it)
            const uint2 mm = fm[lane & (kCT - 1)];
            const int c16 = __popc(mm.x) + __popc(mm.y);
            int incl = c16;
#pragma unroll
            for (int o = 1; o < kCT; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((lane & (kCT - 1)) >= o) incl += v;
            }
            const int tot = __shfl_sync(0xffffffffu, incl, kCT - 1);
            if (count + tot > kCS) {                          // warp-uniform and the same in every warp
                flush();
                count = 0;
            }
            for (int t = warp; t < nt; t += kTW) {
                const int off = count + __shfl_sync(0xffffffffu, incl - c16, t);
                const uint32_t b0 = __shfl_sync(0xffffffffu, mm.x, t), b1 = __shfl_sync(0xffffffffu, mm.y, t);
                if ((b0 | b1) == 0) continue;
                const SrcA<T> A = reinterpret_cast<const SrcA<T>*>(base + (size_t)t * TileBytes<T>::v)[lane];
                const SrcB<T> B = reinterpret_cast<const SrcB<T>*>(base + (size_t)t * TileBytes<T>::v + 32 * sizeof(SrcA<T>))[lane];
                const uint32_t lt = (1u << lane) - 1u;
                if ((b0 >> lane) & 1u) write_entry(off + __popc(b0 & lt), A.x0, A.y0, B.c0, B.s0);
                if ((b1 >> lane) & 1u) write_entry(off + __popc(b0) + __popc(b1 & lt), A.x1, A.y1, B.c1, B.s1);
            }
            count += tot;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        ++it;
    }
    if (stats && n_eval) atomicAdd(stats, n_eval);
}

template <typename T>
__global__ void reduce_groups_kernel(const T* __restrict__ partial, int n_groups, int64_t n_tgt, T f0,
                                     T* __restrict__ frep, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tgt * 2) return;
    T acc = (T)0;
    for (int c = 0; c < n_groups; ++c) acc += partial[(size_t)c * n_tgt * 2 + i];
    acc *= f0;
    frep[i] = accumulate ? frep[i] + acc : acc;
}

template <typename T> size_t tiled_smem_bytes() {
    constexpr int kStages = Stages<T>::n;
    return kStages * ((size_t)kCS * sizeof(Xycs<T>) + (size_t)kCT * sizeof(Tile<T>)) +
           2 * kStages * sizeof(uint64_t) + kStages * sizeof(int4) +
           (size_t)kCS * sizeof(Xycs<T>) + (size_t)kCT * sizeof(Tile<T>) + (size_t)kTB * sizeof(Xycs<T>) +
           (size_t)kTB * 2 * sizeof(T) + (size_t)kLobeBins * sizeof(T) + (size_t)kTB * sizeof(uint32_t) +
           (size_t)kStages * kCT * sizeof(uint2) + 4 * sizeof(int);
}

int g_tiled_ctas[2] = {0, 0};
template <typename T> int tiled_ctas() {
    const int idx = sizeof(T) == 4 ? 0 : 1;
    if (g_tiled_ctas[idx] == 0) {
        int best = 1 << 30;
        const size_t smem = tiled_smem_bytes<T>();
        {
            auto kern = pair_tiled_kernel<T, false>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kTThreads, smem);
            best = nb < best ? nb : best;
        }
        {
            auto kern = pair_tiled_kernel<T, true>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kTThreads, smem);
            best = nb < best ? nb : best;
        }
        g_tiled_ctas[idx] = best < 1 ? 1 : best;
    }
    return g_tiled_ctas[idx];
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Work decomposition: target blocks of kTW * tpw targets; if there are too few blocks to keep every
// CTA slot busy with several items, shrink tpw (down to 4) and then split the chunk range in groups.
struct TiledPlan { int tpw, n_tblocks, n_groups, group_chunks, grid; int64_t n_tiles, n_chunks; };
template <typename T> TiledPlan tiled_plan(int64_t n_src, int64_t n_tgt) {
    static const int env_tpw = env_int("CSF_TILED_TPW", 0), env_groups = env_int("CSF_TILED_GROUPS", 0),
                     env_ipw = env_int("CSF_TILED_ITEMS_PER_SLOT", 4);
    TiledPlan p;
    p.n_tiles = (n_src + kTileS - 1) / kTileS;
    p.n_chunks = (p.n_tiles + kCT - 1) / kCT;
    const int64_t slots = (int64_t)csf_sm_count() * tiled_ctas<T>();
    const int64_t want = (int64_t)env_ipw * slots;
    // 64 targets per block balance the lobe filter's cost (per block) against its selectivity (block
    // radius).  Items must be plentiful -- a few per CTA slot -- because their cost follows the local
    // density of the crowd (a jammed cluster costs several times the mean) and an item is the unit of
    // dynamic scheduling: when there are too few blocks, the chunk range of each block is split over
    // several items (each chunk is still filtered once per block).
    int tpw = 8;
    while (tpw > 1 && (int64_t)kTW * tpw > 2 * n_tgt) tpw >>= 1;      // tiny crowds
    if (env_tpw >= 1 && env_tpw <= kMaxTPW) tpw = env_tpw;
    const int64_t tb = (n_tgt + (int64_t)kTW * tpw - 1) / ((int64_t)kTW * tpw);
    int64_t groups = (want + tb - 1) / tb;
    if (groups > 8) groups = 8;
    if (env_groups >= 1) groups = env_groups;
    if (groups < 1) groups = 1;
    if (groups > kTMaxGroups) groups = kTMaxGroups;
    if (groups > p.n_chunks) groups = p.n_chunks;
    const int64_t gc = (p.n_chunks + groups - 1) / groups;
    groups = (p.n_chunks + gc - 1) / gc;
    p.tpw = tpw;
    p.n_tblocks = (int)tb;
    p.n_groups = (int)groups;
    p.group_chunks = (int)gc;
    const int64_t items = tb * groups;
    p.grid = (int)(items < slots ? items : slots);
    return p;
}

// Rigorous lower bound of the decay rate q/sigma [1/m] of |F| = f_0 exp(-rho q / sigma) over
// s2 = sin^2(psi_i - psi_j) in [0,1] and c = cos(phi) in [-1,1] (vehicle.py:1596-1613), by interval
// arithmetic on a grid of cells.  0 if the field has no positive bound (then nothing is truncated).
double field_min_decay_rate(const CsfFieldParams* fp) {
    const int NS = 256, NC = 512;
    double best = 1e300;
    for (int i = 0; i < NS; ++i) {
        const double s_lo = (double)i / NS, s_hi = (double)(i + 1) / NS;
        const double e_a = fp->e_0 - fp->e_1 * s_lo, e_b = fp->e_0 - fp->e_1 * s_hi;
        const double emax = fmax(fabs(e_a), fabs(e_b));
        const double A_hi = fmax(fp->sigma_0 + fp->sigma_1 * s_lo, fp->sigma_0 + fp->sigma_1 * s_hi);
        const double B_a = fp->sigma_2 + fp->sigma_3 * s_lo, B_b = fp->sigma_2 + fp->sigma_3 * s_hi;
        for (int j = 0; j < NC; ++j) {
            const double c_lo = -1.0 + 2.0 * j / NC, c_hi = -1.0 + 2.0 * (j + 1) / NC;
            const double cmax = fmax(fabs(c_lo), fabs(c_hi));
            const double q2 = 1.0 - emax * emax * cmax * cmax;
            if (!(q2 > 0.0)) return 0.0;
            const double h_lo = sqrt(fmax(0.0, (1.0 - c_hi) * 0.5)), h_hi = sqrt(fmax(0.0, (1.0 - c_lo) * 0.5));
            const double bh = fmin(fmin(B_a * h_lo, B_a * h_hi), fmin(B_b * h_lo, B_b * h_hi));
            const double sig_hi = A_hi - bh;
            if (!(sig_hi > 0.0)) return 0.0;
            best = fmin(best, sqrt(q2) / sig_hi);
        }
    }
    return best < 1e300 ? best : 0.0;
}

// Per c-cell lower bound of the decay rate (min over all heading differences), same interval
// arithmetic as field_min_decay_rate; rate_c[j] covers cos(phi) in [-1 + 2j/NC, -1 + 2(j+1)/NC].
// Returns false if the field has no positive bound somewhere.
bool field_decay_rate_by_angle(const CsfFieldParams* fp, int NC, double* rate_c) {
    const int NS = 256;
    for (int j = 0; j < NC; ++j) rate_c[j] = 1e300;
    for (int i = 0; i < NS; ++i) {
        const double s_lo = (double)i / NS, s_hi = (double)(i + 1) / NS;
        const double e_a = fp->e_0 - fp->e_1 * s_lo, e_b = fp->e_0 - fp->e_1 * s_hi;
        const double emax = fmax(fabs(e_a), fabs(e_b));
        const double A_hi = fmax(fp->sigma_0 + fp->sigma_1 * s_lo, fp->sigma_0 + fp->sigma_1 * s_hi);
        const double B_a = fp->sigma_2 + fp->sigma_3 * s_lo, B_b = fp->sigma_2 + fp->sigma_3 * s_hi;
        for (int j = 0; j < NC; ++j) {
            const double c_lo = -1.0 + 2.0 * j / NC, c_hi = -1.0 + 2.0 * (j + 1) / NC;
            const double cmax = fmax(fabs(c_lo), fabs(c_hi));
            const double q2 = 1.0 - emax * emax * cmax * cmax;
            if (!(q2 > 0.0)) return false;
            const double h_lo = sqrt(fmax(0.0, (1.0 - c_hi) * 0.5)), h_hi = sqrt(fmax(0.0, (1.0 - c_lo) * 0.5));
            const double bh = fmin(fmin(B_a * h_lo, B_a * h_hi), fmin(B_b * h_lo, B_b * h_hi));
            const double sig_hi = A_hi - bh;
            if (!(sig_hi > 0.0)) return false;
            rate_c[j] = fmin(rate_c[j], sqrt(q2) / sig_hi);
        }
    }
    return true;
}

template <typename T> CullConst<T> make_cull(const CsfFieldParams* fp, bool is_f32) {
    CullConst<T> c;
    const double a = fmin(fp->hfov * 0.5, CSF_PI);
    c.ca = (T)cos(a);
    c.sa = (T)sin(a);
    if (a >= CSF_PI) { c.ca = (T)-1; c.sa = (T)0; }
    const T never = (T)(is_f32 ? 3.0e9 : 1e150);          // (payload positions span < 2^31 units)
    c.dmax = never;
    for (int b = 0; b < kLobeBins; ++b) c.lobe[b] = never;
    if (is_f32) {
        // exp(-d * rate) = 2^-cutoff_log2  ->  reach d.  The bounds cost ~1 ms of host time per parameter
        // set: a small cache (several source classes alternate within a step)
        constexpr int NC = 512, kCache = 8;
        struct Entry { CsfFieldParams fp; double rate; double rate_c[NC]; bool ok; bool used; };
        static Entry cache[kCache];
        static int next_slot = 0;
        Entry* hit = nullptr;
        for (int i = 0; i < kCache && !hit; ++i) {
            const Entry& e = cache[i];
            if (e.used && e.fp.e_0 == fp->e_0 && e.fp.e_1 == fp->e_1 && e.fp.sigma_0 == fp->sigma_0 &&
                e.fp.sigma_1 == fp->sigma_1 && e.fp.sigma_2 == fp->sigma_2 && e.fp.sigma_3 == fp->sigma_3)
                hit = &cache[i];
        }
        if (!hit) {
            hit = &cache[next_slot];
            next_slot = (next_slot + 1) % kCache;
            hit->fp = *fp;
            hit->rate = field_min_decay_rate(fp);
            hit->ok = field_decay_rate_by_angle(fp, NC, hit->rate_c);
            hit->used = true;
        }
        const double cached_rate = hit->rate;
        const double* cached_rate_c = hit->rate_c;
        const bool cached_ok = hit->ok;
        const double bits = fp->cutoff_log2 > 0.0 ? fp->cutoff_log2 : 40.0;
        const double L = bits * 0.6931471805599453 * 1.0001 / fp->q_scale;
        if (cached_rate > 0.0 && L / cached_rate < 3.0e9) {
            c.dmax = (T)(L / cached_rate);
            if (cached_ok) {
                // lobe[b] = max reach over all cells with cos(phi) below the bin's upper edge
                double run = 0.0;
                const int per = NC / kLobeBins;
                for (int b = 0; b < kLobeBins; ++b) {
                    for (int j = b * per; j < (b + 1) * per; ++j) run = fmax(run, L / cached_rate_c[j]);
                    c.lobe[b] = (T)fmin(run, (double)c.dmax);
                }
            } else {
                for (int b = 0; b < kLobeBins; ++b) c.lobe[b] = c.dmax;
            }
        }
    }
    return c;
}

template <typename T>
int tile_sources(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles, cudaStream_t st) {
    if (n <= 0) return 0;
    const int64_t n_tiles = (n + kTileS - 1) / kTileS, n_chunks = (n_tiles + kCT - 1) / kCT;
    tile_sources_kernel<T><<<(unsigned)n_chunks, kCT * 32, 0, st>>>(
        (const Xycs<T>*)xycs, n, perm, (unsigned char*)sorted, (Tile<T>*)tiles, n_tiles);
    CSF_CHECK_LAUNCH("tile_sources_kernel");
    return 0;
}

constexpr size_t kWsHeader = 256;   // item counter
template <typename T> size_t tiled_ws_blocks_bytes(const TiledPlan& pl) {
    return (((size_t)pl.n_tblocks * sizeof(Tile<T>)) + 255) / 256 * 256;
}

template <typename T>
int pair_tiled(const void* sorted, const void* tiles, int64_t n_src, const void* tgt, const int64_t* tgt_perm,
               int64_t n_tgt, const CsfFieldParams* fp, T* frep, int accumulate, void* ws, size_t wsb,
               unsigned long long* stats, cudaStream_t st) {
    if (n_tgt <= 0) return 0;
    if (n_src <= 0 || fp->f_0 == 0.0) {
        if (!accumulate) cudaMemsetAsync(frep, 0, sizeof(T) * 2 * n_tgt, st);
        return 0;
    }
    const TiledPlan pl = tiled_plan<T>(n_src, n_tgt);
    const size_t off_partial = kWsHeader + tiled_ws_blocks_bytes<T>(pl);
    const size_t need = off_partial + (size_t)pl.n_groups * n_tgt * 2 * sizeof(T);
    if (ws == nullptr || wsb < need) {
        csf_set_error("csf_pair_forces_tiled: workspace too small", cudaErrorInvalidValue);
        return -(int)cudaErrorInvalidValue;
    }
    const PairConst<T> k = make_const<T>(fp, sizeof(T) == 4);
    const CullConst<T> cc = make_cull<T>(fp, sizeof(T) == 4);
    const size_t smem = tiled_smem_bytes<T>();
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
    Tile<T>* tblocks = reinterpret_cast<Tile<T>*>((unsigned char*)ws + kWsHeader);
    T* partial = reinterpret_cast<T*>((unsigned char*)ws + off_partial);
    block_bounds_kernel<T><<<(unsigned)(((int64_t)pl.n_tblocks * 32 + 127) / 128), 128, 0, st>>>(
        (const Xycs<T>*)tgt, tgt_perm, n_tgt, kTW * pl.tpw, tblocks, pl.n_tblocks, counter);
    CSF_CHECK_LAUNCH("block_bounds_kernel");
    tiled_ctas<T>();   // sets the dynamic shared-memory attribute
    if (fp->p2r)
        pair_tiled_kernel<T, true><<<pl.grid, kTThreads, smem, st>>>(
            (const unsigned char*)sorted, (const Tile<T>*)tiles, pl.n_tiles, (const Xycs<T>*)tgt, tgt_perm, n_tgt,
            tblocks, k, cc, partial, pl.tpw, pl.group_chunks, pl.n_groups, pl.n_tblocks, counter, stats);
    else
        pair_tiled_kernel<T, false><<<pl.grid, kTThreads, smem, st>>>(
            (const unsigned char*)sorted, (const Tile<T>*)tiles, pl.n_tiles, (const Xycs<T>*)tgt, tgt_perm, n_tgt,
            tblocks, k, cc, partial, pl.tpw, pl.group_chunks, pl.n_groups, pl.n_tblocks, counter, stats);
    CSF_CHECK_LAUNCH("pair_tiled_kernel");
    const int64_t n2 = n_tgt * 2;
    reduce_groups_kernel<T><<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(partial, pl.n_groups, n_tgt, (T)fp->f_0, frep,
                                                                           accumulate);
    CSF_CHECK_LAUNCH("reduce_groups_kernel");
    return 0;
}

}  // namespace

extern "C" {

int64_t csf_tiled_padded_sources(int64_t n_src) { return ((n_src + kTileS - 1) / kTileS) * kTileS; }
int64_t csf_tiled_num_tiles(int64_t n_src) {
    const int64_t t = (n_src + kTileS - 1) / kTileS;
    return t + (t + kCT - 1) / kCT;          // tile records followed by chunk records
}
int csf_tiled_tile_bytes(int elem_bytes) { return elem_bytes == 4 ? (int)sizeof(Tile<float>) : (int)sizeof(Tile<double>); }
size_t csf_pair_tiled_workspace_bytes(int64_t n_src, int64_t n_tgt, int elem_bytes) {
    if (n_src <= 0 || n_tgt <= 0) return 0;
    if (elem_bytes == 4) {
        const TiledPlan pl = tiled_plan<float>(n_src, n_tgt);
        return kWsHeader + tiled_ws_blocks_bytes<float>(pl) + (size_t)pl.n_groups * (size_t)n_tgt * 2 * 4;
    }
    const TiledPlan pl = tiled_plan<double>(n_src, n_tgt);
    return kWsHeader + tiled_ws_blocks_bytes<double>(pl) + (size_t)pl.n_groups * (size_t)n_tgt * 2 * 8;
}
double csf_field_cutoff_distance(const CsfFieldParams* fp) {
    const double rate = field_min_decay_rate(fp);
    const double bits = fp->cutoff_log2 > 0.0 ? fp->cutoff_log2 : 40.0;
    return rate > 0.0 ? bits * 0.6931471805599453 / rate : INFINITY;
}
int csf_field_reach_table(const CsfFieldParams* fp, int n_bins, double* reach_m) {
    // host-only: the reach table of the lobe filter in metres (bin b covers cos(phi) in
    // [-1 + 2b/n_bins, -1 + 2(b+1)/n_bins]); n_bins must be kLobeBins
    if (n_bins != kLobeBins) return -1;
    CsfFieldParams q = *fp;
    if (!(q.q_scale > 0.0)) q.q_scale = 1.0;
    const CullConst<float> c = make_cull<float>(&q, true);
    for (int b = 0; b < kLobeBins; ++b) reach_m[b] = c.lobe[b] >= 2.9e9f ? INFINITY : (double)c.lobe[b] * q.q_scale;
    return 0;
}
int csf_morton_keys_f32(const void* xycs, int64_t n, double x0, double y0, double cell, int64_t* keys, csf_stream_t st) {
    if (n <= 0) return 0;
    morton_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>((const Xycs<float>*)xycs, n, x0, y0,
                                                                                  1.0 / cell, nullptr, keys);
    CSF_CHECK_LAUNCH("morton_kernel");
    return 0;
}
int csf_morton_keys_f64(const void* xycs, int64_t n, double x0, double y0, double cell, int64_t* keys, csf_stream_t st) {
    if (n <= 0) return 0;
    morton_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>((const Xycs<double>*)xycs, n, x0,
                                                                                   y0, 1.0 / cell, nullptr, keys);
    CSF_CHECK_LAUNCH("morton_kernel");
    return 0;
}
int csf_spatial_keys_f32(const void* xycs, int64_t n, const double* box_dev, int64_t* keys, csf_stream_t st) {
    if (n <= 0) return 0;
    morton_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>((const Xycs<float>*)xycs, n, 0.0, 0.0,
                                                                                  1.0, box_dev, keys);
    CSF_CHECK_LAUNCH("morton_kernel");
    return 0;
}
int csf_spatial_keys_f64(const void* xycs, int64_t n, const double* box_dev, int64_t* keys, csf_stream_t st) {
    if (n <= 0) return 0;
    morton_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>((const Xycs<double>*)xycs, n, 0.0,
                                                                                   0.0, 1.0, box_dev, keys);
    CSF_CHECK_LAUNCH("morton_kernel");
    return 0;
}
int csf_tile_sources_f32(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles, csf_stream_t st) {
    return tile_sources<float>(xycs, n, perm, sorted, tiles, (cudaStream_t)st);
}
int csf_tile_sources_f64(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles, csf_stream_t st) {
    return tile_sources<double>(xycs, n, perm, sorted, tiles, (cudaStream_t)st);
}
int csf_pair_forces_tiled_f32(const void* sorted, const void* tiles, int64_t n_src, const void* tgt,
                              const int64_t* tgt_perm, int64_t n_tgt, const CsfFieldParams* fp, float* frep,
                              int accumulate, void* ws, size_t wsb, unsigned long long* stats, csf_stream_t st) {
    return pair_tiled<float>(sorted, tiles, n_src, tgt, tgt_perm, n_tgt, fp, frep, accumulate, ws, wsb, stats,
                             (cudaStream_t)st);
}
int csf_pair_forces_tiled_f64(const void* sorted, const void* tiles, int64_t n_src, const void* tgt,
                              const int64_t* tgt_perm, int64_t n_tgt, const CsfFieldParams* fp, double* frep,
                              int accumulate, void* ws, size_t wsb, unsigned long long* stats, csf_stream_t st) {
    return pair_tiled<double>(sorted, tiles, n_src, tgt, tgt_perm, n_tgt, fp, frep, accumulate, ws, wsb, stats,
                              (cudaStream_t)st);
}

}  // extern "C"
