// K1 (tiled): all-pairs repulsive force with hierarchical culling, warp-specialised.
//
// The masked sum  Frep_j = sum_i mask(i,j) F(i -> j)  (reference intersection.py:788-843,
// vehicle.py:1560-1648) has three families of exact or negligible zeros:
//   * ~2/3 of the ordered pairs are masked by the *target's* field of view
//     (intersection.py:733-736; hfov = 2 pi / 3);
//   * |F(i -> j)| = f_0 exp(-rho q / sigma) (vehicle.py:1613-1648): beyond rho = d_cut every
//     contribution is below 2^-cutoff_log2 f_0 (f32 build only; d_cut ~ 160 m with the default
//     parameters; the f64 verification build never truncates);
//   * the reach of a source is strongly anisotropic (160 m ahead of it, 55 m from 90 deg on).
// Sources live in a Hilbert-ordered copy cut into tiles of 64 and chunks of 16 tiles, each with a
// bounding circle; targets are visited in the same order in blocks of 64 with a bounding circle.
// A work item is (target block x chunk group).  A CTA is three kinds of warps that only meet
// through mbarriers -- no CTA-wide barrier anywhere:
//   producer (1 warp)  fetches items from an atomic counter, tests chunk and tile circles against
//                      the block circle and streams the tiles that can be within d_cut of the block
//                      into a ring of shared-memory stages (8 tiles each) with 1-D TMA bulk copies;
//   filter   (2 warps; 4 in the wide shape) take the stages, drop every source whose field cannot reach the block circle
//                      (lobe test, packed FP32x2) and append the survivors, in stream order, to one
//                      of two survivor buffers (32 dynamic tiles of 64, with bounding circles); they
//                      also load the block's targets;
//   evaluate (11 warps; 22 in the wide shape) take the block's targets one at a time from a shared
//                      counter: per survivor buffer a warp tests the target's view cone against the 32
//                      circles (one ballot), evaluates the surviving tiles two sources per lane
//                      (pair_eval2, which still applies the exact per-pair mask), reduces both force
//                      components with one butterfly and adds them to the target's accumulator.  The
//                      buffer is handed back when all evaluate warps are done with it.
// While the evaluate warps work on one buffer the filter warps fill the other, and the producer is
// stages ahead of both.  Culled pairs contribute exactly 0 (mask) or < 2^-cutoff_log2 f_0 (f32), so
// the result equals the dense kernel's up to the order of summation and that bound.  Every sum has
// a fixed order (stream order of the survivors, one warp per (target, buffer)): deterministic, no
// float atomics; partial sums per chunk group are reduced in fixed order.
#include "csf_common.cuh"
#include "csf_pair_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "csf_peer.cuh"

namespace {

// Two shapes of CTA, same code (template parameters FW / EW = filter / evaluate warps):
//   narrow  1 + 2 + 11 warps, two CTAs per SM   -- several items per CTA slot: the CTAs of an SM fill each
//                                                  other's gaps at item boundaries.  (Measured on B200,
//                                                  N = 65,536 / a half / an eighth of it, K1 in us:
//                                                  1+2+6 x3: 305 / 186 / 70;  1+2+11 x2: 306 / 174 / 69;
//                                                  1+4+11 x2: 303 / 174 / 68;  1+2+12 x2 (64 registers, spills):
//                                                  314 / 178 / 70;  1+1+6 x3: 324 / 205 / 69 -- 22 evaluate
//                                                  warps per SM at 72 registers instead of 18 change little:
//                                                  the kernel sits on a plateau that the shape does not move);
//   wide    1 + 4 + 22 warps, one CTA per SM    -- few items (a small crowd, a rank's shard of a crowd): a
//                                                  CTA that has an SM to itself puts all of the SM's
//                                                  evaluate warps on ONE item, whose latency is what
//                                                  bounds such a launch.
#ifndef CSF_TILED_MINB
#define CSF_TILED_MINB 2
#endif
#ifndef CSF_NARROW_FW
#define CSF_NARROW_FW 2
#endif
#ifndef CSF_NARROW_EW
#define CSF_NARROW_EW 11
#endif
constexpr int kNarrowFW = CSF_NARROW_FW, kNarrowEW = CSF_NARROW_EW, kWideFW = 4, kWideEW = 22;
constexpr int kBT = 64;                  // targets per block
constexpr int kTileS = 64;               // sources per tile (2 per lane)
constexpr int kCT = 16;                  // tiles per chunk of the sorted copy
constexpr int kCS = kCT * kTileS;        // sources per chunk
// Dynamic tiles per survivor buffer: 16 (two targets' view cones per cull ballot) or 32 (one target per
// ballot, half as many (target, buffer) units for a large item -- their fixed cost, cull + reductions, is a
// fifth of the evaluate warps' instructions).  f64 buffers are twice the size: 16 there.
#ifndef CSF_TILED_DT
#define CSF_TILED_DT 32
#endif
template <typename T> struct DynTiles { static constexpr int n = sizeof(T) == 4 ? CSF_TILED_DT : 16; };
static_assert(CSF_TILED_DT == 16 || CSF_TILED_DT == 32, "dynamic tiles per survivor buffer: 16 or 32");
constexpr int kST = 8;                   // tiles per shared-memory stage
constexpr int kTMaxGroups = 64;
#ifndef CSF_TILED_ITEMS_IN_FLIGHT
#define CSF_TILED_ITEMS_IN_FLIGHT 2
#endif
constexpr int kItemsInFlight = CSF_TILED_ITEMS_IN_FLIGHT;   // items a CTA may hold (claimed, not yet evaluated)
template <typename T> struct Stages { static constexpr int n = 4; };
static_assert(kST % kNarrowFW == 0 && kST % kWideFW == 0, "filter warps split stages evenly");
// Capacity of the k-th survivor buffer of an item (slow start): the evaluate warps get their first
// buffer of a new item after 256 survivors instead of 1024 -- the bubble at every item boundary is a
// quarter as long -- and steady-state buffers are large, so that the fixed cost per (target, buffer)
// stays small.  A fixed function of k: the order of every sum stays fixed.
#ifndef CSF_TILED_SLOW_START
#define CSF_TILED_SLOW_START 1
#endif
// (tiles; narrow shape / wide shape.  Measured with 32-tile buffers, step in us at N = 65,536 / 1 M / a 1/8
// shard (wide): 4, 8: 298 / 4602 / 73;  8, 16: 285 / 4281 / 75;  4, 16: 287 / 4399 / 76;  none: 289 / 4333 / 82;
// 16-tile buffers with 4, 8: 299 / 4535 / 78.)
#ifndef CSF_TILED_CAP0
#define CSF_TILED_CAP0 8
#endif
#ifndef CSF_TILED_CAP1
#define CSF_TILED_CAP1 16
#endif
#ifndef CSF_TILED_WIDE_CAP0
#define CSF_TILED_WIDE_CAP0 4
#endif
#ifndef CSF_TILED_WIDE_CAP1
#define CSF_TILED_WIDE_CAP1 8
#endif
template <int DT, bool WIDE> __device__ __forceinline__ int buffer_cap(int k) {
    if (!CSF_TILED_SLOW_START) return DT * kTileS;
    const int c0 = (WIDE ? CSF_TILED_WIDE_CAP0 : CSF_TILED_CAP0) < DT ? (WIDE ? CSF_TILED_WIDE_CAP0 : CSF_TILED_CAP0) : DT;
    const int c1 = (WIDE ? CSF_TILED_WIDE_CAP1 : CSF_TILED_CAP1) < DT ? (WIDE ? CSF_TILED_WIDE_CAP1 : CSF_TILED_CAP1) : DT;
    return k == 0 ? c0 * kTileS : (k == 1 ? c1 * kTileS : DT * kTileS);
}

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
// v0.1 Bicycle field (FIELD 1; the sources' headings arrive scaled by their eccentricity): either precision,
// one scalar evaluation per source -- a legacy field, not on the benchmarked path
template <typename T, bool P2R> struct TileAccBike {
    T x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    __device__ __forceinline__ void eval(const SrcA<T>& A, const SrcB<T>& B, const Tgt<T>& tg, const PairConst<T>& k) {
        pair_eval_bike<T, P2R>(src0(A, B), tg, k, x0, y0);
        pair_eval_bike<T, P2R>(src1(A, B), tg, k, x1, y1);
    }
    __device__ __forceinline__ void merge(const TileAccBike& o) { x0 += o.x0; y0 += o.y0; x1 += o.x1; y1 += o.y1; }
    __device__ __forceinline__ void total(T& sx, T& sy) const { sx = x0 + x1; sy = y0 + y1; }
};
template <typename T, bool P2R, int FIELD> struct TileAccSel { typedef TileAccBike<T, P2R> type; };
template <bool P2R> struct TileAccSel<float, P2R, 0> { typedef TileAcc32<P2R> type; };
template <bool P2R> struct TileAccSel<double, P2R, 0> { typedef TileAcc64<P2R> type; };

// ---- bounding circles ---------------------------------------------------------------------------
__device__ __forceinline__ void pad_entry(Xycs<float>& e) { e.xq = 1 << 30; e.yq = 1 << 30; e.c = 1.f; e.s = 0.f; }
__device__ __forceinline__ void pad_entry(Xycs<double>& e) { e.x = 1e150; e.y = 1e150; e.c = 1.0; e.s = 0.0; }

// running bounding box of payload positions (integer for the Q-format payload: exact)
template <typename T> struct BBox;
template <> struct BBox<float> {
    int xmin = INT32_MAX, xmax = INT32_MIN, ymin = INT32_MAX, ymax = INT32_MIN;
    __device__ __forceinline__ void add(int x, int y) { xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y); }
    __device__ __forceinline__ void add(const Xycs<float>& e) { add(e.xq, e.yq); }
    __device__ __forceinline__ void warp_reduce() {      // REDUX.MIN/MAX: one instruction per bound
        xmin = __reduce_min_sync(0xffffffffu, xmin);
        xmax = __reduce_max_sync(0xffffffffu, xmax);
        ymin = __reduce_min_sync(0xffffffffu, ymin);
        ymax = __reduce_max_sync(0xffffffffu, ymax);
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
// speed: sources with the v0.1 Bicycle field (nullptr: TwoD field) -- their entries of the sorted copy carry the
// heading scaled by the eccentricity e = min((v / v_max)^0.1, 0.7) (vehicle.py:1054-1064), padding e = 0.
template <typename T> __device__ __forceinline__ void scale_heading(Xycs<T>& a, bool valid, const T* speed, int64_t i, T v_max) {
    const T e = valid ? (T)fmin(pow((double)speed[i] / (double)v_max, 0.1), 0.7) : (T)0;
    a.c *= e;
    a.s *= e;
}
template <typename T>
__global__ void __launch_bounds__(kCT * 32)
tile_sources_kernel(const Xycs<T>* __restrict__ xycs, int64_t n, const int64_t* __restrict__ perm,
                    unsigned char* __restrict__ sorted, Tile<T>* __restrict__ tiles, int64_t n_tiles,
                    const T* __restrict__ speed, T v_max) {
    __shared__ __align__(16) unsigned char boxes_raw[kCT * sizeof(BBox<T>)];
    BBox<T>* boxes = reinterpret_cast<BBox<T>*>(boxes_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t c = blockIdx.x, t = c * kCT + w;
    BBox<T> bb;
    if (t < n_tiles) {
        const int64_t i0 = t * kTileS + lane, i1 = i0 + 32;
        Xycs<T> a, b;
        const bool va = i0 < n, vb = i1 < n;
        const int64_t j0 = va ? (perm ? perm[i0] : i0) : 0, j1 = vb ? (perm ? perm[i1] : i1) : 0;
        if (va) a = xycs[j0]; else pad_entry(a);
        if (vb) b = xycs[j1]; else pad_entry(b);
        if (speed) {
            scale_heading(a, va, speed, j0, v_max);
            scale_heading(b, vb, speed, j1, v_max);
        }
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

// Tile build and target-block bounds in ONE launch (the step's first kernel): CTAs [0, n_chunks) do what
// tile_sources_kernel does, the rest what block_bounds_kernel does, 16 blocks per CTA.  On a sharded crowd
// every CTA first waits until the peers' payload pushes of the previous step have landed (csf_peer.cu).
template <typename T>
__global__ void __launch_bounds__(kCT * 32)
tiled_prepare_kernel(const Xycs<T>* __restrict__ xycs, int64_t n, const int64_t* __restrict__ perm,
                     unsigned char* __restrict__ sorted, Tile<T>* __restrict__ tiles, int64_t n_tiles, int64_t n_chunks,
                     const Xycs<T>* __restrict__ tgt, const int64_t* __restrict__ tgt_perm, int64_t n_tgt,
                     Tile<T>* __restrict__ blocks, int64_t n_blocks, unsigned int* __restrict__ item_counter,
                     CsfPeerComm comm) {
    // a pair kernel launched as a programmatic dependent may bring its CTAs up while this kernel runs (it reads
    // nothing of ours before its griddepcontrol.wait)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (comm.world > 1) {
        csf_peer_wait_all(comm);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *item_counter = 0;   // the pair kernel's dynamic item counter
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if ((int64_t)blockIdx.x >= n_chunks) {
        const int64_t b = ((int64_t)blockIdx.x - n_chunks) * kCT + w;
        if (b >= n_blocks) return;
        BBox<T> bb;
        const int64_t i_end = min(n_tgt, (b + 1) * (int64_t)kBT);
        for (int64_t i = b * (int64_t)kBT + lane; i < i_end; i += 32) bb.add(tgt[tgt_perm ? tgt_perm[i] : i]);
        bb.warp_reduce();
        if (lane == 0) blocks[b] = bb.circle(i_end - b * (int64_t)kBT);
        return;
    }
    __shared__ __align__(16) unsigned char boxes_raw[kCT * sizeof(BBox<T>)];
    BBox<T>* boxes = reinterpret_cast<BBox<T>*>(boxes_raw);
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

// v0.1 Bicycle field: the level set b = const of a source is an ellipse with the source in one focus, eccentricity
// e and the long axis along the heading -- reach(phi0) = K sqrt(1 - e^2) / (1 - e cos phi0), K = lobe[0] (payload
// units; b = K / p_decay is where |F| has fallen to 2^-cutoff_log2 p_0 / p_decay).  (c, s) = heading scaled by e.
// w = max of e cos(phi) over the directions under which the source sees the block circle, d - R_b = lower bound
// of the distance: the source cannot matter if (d - R_b) (1 - w) > K sqrt(1 - e^2).
template <typename T, typename P>
__device__ __forceinline__ bool bike_reaches(const Tile<T>& blk, P x, P y, T c, T s, T K, T tiny) {
    T dx, dy;
    block_delta(blk, x, y, dx, dy);
    const T d2 = fma(dy, dy, fma(dx, dx, tiny));
    const T rinv = M<T>::rsqrt(d2);
    const T d = d2 * rinv;
    const T Rb = (T)blk.R;
    const T e2 = fmin(fma(s, s, c * c), (T)0.5);          // e <= 0.7
    const T e = M<T>::sqrt(e2);
    const T cphi = fma(dy, s, dx * c) * rinv;             // e cos(phi0)
    const T sphi = fabs(fma(dy, c, -(dx * s))) * rinv;    // e |sin(phi0)|
    const T sdel = fmin(Rb * rinv, (T)1);
    const T cdel = M<T>::sqrt(fmax(fma(-sdel, sdel, (T)1), (T)0));
    T w = fma(cphi, cdel, sphi * sdel);                   // e cos(|phi0| - delta)
    w = (cphi >= e * cdel) ? e : w;                       // the circle straddles the heading direction
    w = fmin(w, e);
    return (d - Rb) * ((T)1 - w) <= K * M<T>::sqrt((T)1 - e2) * (T)1.0001;
}

// named barrier 1 among the filter warps only
template <int FW> __device__ __forceinline__ void filter_barrier() {
    if (FW > 1) asm volatile("bar.sync 1, %0;" ::"n"(FW * 32) : "memory");
    else __syncwarp();
}

enum { BUF_LAST = 1, BUF_EXIT = 2 };

// Optional cycle accounting per warp role (build with -DCSF_TILED_PROF; tools/k1_roles.py prints it):
// stats[1..] += cycles {evaluate: total, waiting for a buffer; filter: total, waiting for a stage,
// waiting for a free slot; producer: total, waiting for a stage slot}, buffers, stages, (target, buffer)
// units evaluated.
// suspend-time hints of the mbarrier waits (ns): hand-offs between the warp roles
#ifndef CSF_TILED_HINT_NS
#define CSF_TILED_HINT_NS 1000u
#endif
#ifndef CSF_TILED_PRODUCER_HINT_NS
#define CSF_TILED_PRODUCER_HINT_NS 4000u
#endif
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifdef CSF_TILED_PROF
#define PROF_T0(v) const long long v = clock64()
#define PROF_ADD(acc, v) acc += clock64() - v
#define PROF_INC(acc) acc += 1
#else
#define PROF_T0(v)
#define PROF_ADD(acc, v)
#define PROF_INC(acc)
#endif

// ---- the tiled pair kernel ---------------------------------------------------------------------------
// item -> (target block tb = item % n_tblocks, part cg = item / n_tblocks of the block's near chunks);
// partial[cg][target][2].
// Stage protocol (producer -> filter warps): hdr[stage] = {item, n}: n = 1..kST tiles of the sorted copy
// with their circle records; n == -1 closes the item, n == -2 ends the kernel.
// Buffer protocol (filter -> evaluate warps): bdesc[slot] = {item, dynamic tiles, flags, target set}:
// BUF_LAST = last buffer of the item (the evaluate warps write their sums), BUF_EXIT = leave.
// FIELD 0: TwoD field (vehicle.py:1560-1648); FIELD 1: v0.1 Bicycle field (vehicle.py:1054-1147; bike_reaches, pair_eval_bike).
template <typename T, bool P2R, int FW, int EW, int FIELD = 0>
__global__ void __launch_bounds__((1 + FW + EW) * 32, (sizeof(T) == 4 && EW == kNarrowEW) ? CSF_TILED_MINB : 1)
pair_tiled_kernel(const unsigned char* __restrict__ sorted, const Tile<T>* __restrict__ tiles, int64_t n_tiles,
                  const Xycs<T>* __restrict__ tgt, const int64_t* __restrict__ tgt_perm, int64_t n_tgt,
                  const Tile<T>* __restrict__ tblocks, PairConst<T> k, CullConst<T> cc, T* __restrict__ partial,
                  int n_groups, int n_tblocks, unsigned int* __restrict__ counter,
                  const unsigned int* __restrict__ item_order, unsigned int* __restrict__ item_cost,
                  unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kStages = Stages<T>::n;
    constexpr size_t kTileB = TileBytes<T>::v;
    unsigned char* stage_src = smem_raw;                                            // [kStages][kST] tiles
    constexpr int kDT = DynTiles<T>::n;                                             // dynamic tiles per survivor buffer
    constexpr int kDS = kDT * kTileS;                                               // survivors per buffer
    unsigned char* sbuf = stage_src + (size_t)kStages * kST * kTileB;               // [2][kDT] survivor tiles
    Tile<T>* srec = reinterpret_cast<Tile<T>*>(sbuf + (size_t)2 * kDT * kTileB);    // [kStages][kST] circles of the staged tiles
    Tile<T>* dtile = srec + kStages * kST;                                          // [2][kDT] circles of the survivor tiles
    Xycs<T>* btgt = reinterpret_cast<Xycs<T>*>(dtile + 2 * kDT);                    // [2][kBT] targets, by heading rank
    long long* bj = reinterpret_cast<long long*>(btgt + 2 * kBT);                   // [2][kBT] their indices
    uint64_t* full = reinterpret_cast<uint64_t*>(bj + 2 * kBT);                     // [kStages]
    uint64_t* empty = full + kStages;                                               // [kStages]
    uint64_t* ready = empty + kStages;                                              // [2]
    uint64_t* freeb = ready + 2;                                                    // [2]
    int4* hdr = reinterpret_cast<int4*>(freeb + 2);                                 // [kStages] {item, n, -, -}
    int4* bdesc = hdr + kStages;                                                    // [2] {item, n_dt, flags, target set}
    uint2* fmask = reinterpret_cast<uint2*>(bdesc + 2);                             // [kStages][kST] filter ballots
    T* lobe = reinterpret_cast<T*>(fmask + kStages * kST);                          // [kLobeBins] reach table
    T* bacc = lobe + kLobeBins;                                                     // [2 items][2 buffer parities][kBT][2] sums
    int* nextq = reinterpret_cast<int*>(bacc + 2 * 2 * kBT * 2);                    // [2] next target of the buffer in the slot
    int* last_seen = nextq + 2;                                                     // evaluate warps that have left an item's last buffer, summed over items

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], FW);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&ready[s], 1);
            mbar_init(&freeb[s], EW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kLobeBins; i += blockDim.x) lobe[i] = cc.lobe[i];
    for (int i = threadIdx.x; i < 2 * 2 * kBT * 2; i += blockDim.x) bacc[i] = (T)0;
    if (threadIdx.x == 0) *last_seen = 0;
    __syncthreads();
    // everything above touches shared memory and kernel parameters only; from here on the kernel reads what the
    // preceding kernel (tile build, block bounds, item counter) wrote.  A no-op unless launched with CSF_TILED_PDL.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const unsigned int n_items = (unsigned int)n_tblocks * (unsigned int)n_groups;
    const int64_t n_chunks = (n_tiles + kCT - 1) / kCT;
    const Tile<T>* chunks = tiles + n_tiles;

    if (warp == 0) {
        // ===== producer warp: fetch items, cull chunks and tiles against the target block, stream the rest =====
        uint32_t it = 0;
        int fill = 0;                                   // tiles in the open stage
        long long pr_wait = 0;
        PROF_T0(pr_t0);
        auto close_stage = [&](int item, int n) {
            __syncwarp();                               // the circle records of every lane have been written
            if (lane == 0) {
                const int stage = it % kStages;
                hdr[stage] = make_int4(item, n, 0, 0);
                mbar_arrive(&full[stage]);
            }
            ++it;
            fill = 0;
        };
        auto open_stage = [&]() {                       // wait until the filter warps have released the stage
            PROF_T0(w0);
            if (lane == 0) mbar_wait_backoff(&empty[it % kStages], ((it / kStages) & 1) ^ 1, CSF_TILED_PRODUCER_HINT_NS);
            __syncwarp();
            PROF_ADD(pr_wait, w0);
        };
        int fetched = 0;                                // items this CTA has claimed
        for (;;) {
            // An item is claimed only when the last item but one has been evaluated completely: one item with
            // the evaluate warps, one with the filter warps.  Without the limit the producer -- up to kStages
            // stages ahead -- hoards the small items of a small crowd or shard while other CTAs run dry.
            if (lane == 0 && fetched >= kItemsInFlight) {
                const int want = EW * (fetched - kItemsInFlight + 1);
                while (*reinterpret_cast<volatile int*>(last_seen) < want) __nanosleep(200);
            }
            __syncwarp();
            unsigned int item = 0;
            if (lane == 0) item = atomicAdd(counter, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= n_items) {
                // this CTA takes no further items: once every CTA has said so (or exited), a kernel launched
                // as a programmatic dependent of this one (the step's per-agent kernel) may start on the SMs
                // that are running dry; whatever it needs from this launch it reads behind griddepcontrol.wait
                asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                break;
            }
            ++fetched;
            if (item_order) item = item_order[item];        // heaviest items first (previous step's costs)
#ifdef CSF_TILED_PROF
            if (stats && lane == 0) { stats[16 + 4 * item] = globaltimer_ns(); stats[16 + 4 * item + 3] = blockIdx.x; }
#endif
            const int tb = (int)(item % (unsigned int)n_tblocks), cg = (int)(item / (unsigned int)n_tblocks);
            const Tile<T> tbr = tblocks[tb];
            // The chunks that can be within d_cut of the block, in order, are dealt out evenly to the block's
            // n_groups items: item cg streams entries [lo, hi) of that list.  (Splitting the chunk INDEX
            // range instead would give one item nearly all of the work: the near chunks of a block are
            // neighbours on the Hilbert curve.)
            int lo = 0, hi = 0x7fffffff;
            if (n_groups > 1) {
                int n_near = 0;
                for (int64_t c0 = 0; c0 < n_chunks; c0 += 32) {
                    const int64_t c = c0 + lane;
                    const bool near = c < n_chunks && circles_near(chunks[c], tbr, cc.dmax);
                    n_near += __popc(__ballot_sync(0xffffffffu, near));
                }
                lo = (int)(((long long)n_near * cg) / n_groups);
                hi = (int)(((long long)n_near * (cg + 1)) / n_groups);
            }
            int seen = 0;                               // near chunks passed so far
            for (int64_t c0 = 0; c0 < n_chunks && seen < hi; c0 += 32) {
                const int64_t c = c0 + lane;
                bool near = false;
                if (c < n_chunks) near = circles_near(chunks[c], tbr, cc.dmax);
                uint32_t m = __ballot_sync(0xffffffffu, near);
                while (m) {
                    const int64_t ch = c0 + (__ffs(m) - 1);
                    m &= m - 1;
                    const int ord = seen++;
                    if (ord < lo) continue;
                    if (ord >= hi) break;
                    const int64_t t0 = ch * kCT;
                    Tile<T> rec;
                    bool tn = false;
                    if (lane < kCT && t0 + lane < n_tiles) {
                        rec = tiles[t0 + lane];
                        tn = circles_near(rec, tbr, cc.dmax);
                    }
                    uint32_t tm = __ballot_sync(0xffffffffu, tn);
                    while (tm) {
                        if (fill == 0) open_stage();
                        const int stage = it % kStages;
                        const int room = kST - fill;
                        const int rank = __popc(tm & ((1u << lane) - 1u));
                        const bool mine = ((tm >> lane) & 1u) && rank < room;
                        const int take = min(__popc(tm), room);
                        if (lane == 0) mbar_expect_tx_only(&full[stage], (uint32_t)take * (uint32_t)kTileB);
                        __syncwarp();
                        if (mine) {
                            const int slot = stage * kST + fill + rank;
                            srec[slot] = rec;
                            tma_bulk_g2s(stage_src + (size_t)slot * kTileB, sorted + (size_t)(t0 + lane) * kTileB,
                                         (uint32_t)kTileB, &full[stage]);
                        }
                        tm &= ~__ballot_sync(0xffffffffu, mine);
                        fill += take;
                        if (fill == kST) close_stage((int)item, kST);
                    }
                }
            }
            if (fill > 0) close_stage((int)item, fill);
            open_stage();
            close_stage((int)item, -1);                 // close the item
        }
        open_stage();
        close_stage(0, -2);
#ifdef CSF_TILED_PROF
        if (stats && lane == 0) {
            atomicAdd(stats + 6, (unsigned long long)(clock64() - pr_t0));
            atomicAdd(stats + 7, (unsigned long long)pr_wait);
        }
#endif
        (void)pr_wait;
        return;
    }

    typedef decltype(SrcA<T>().x0) P;    // payload position type

    if (warp <= FW) {
        // ===== filter warps =====
        // Per stage: (1) every source of the staged tiles is tested against the target block's circle
        // with the lobe test (lobe_reaches2): most sources near the block cannot matter for any of its
        // targets, because a source's field reaches far only in a narrow range of directions;
        // (2) the survivors are appended, in stream order, to the open survivor buffer (same tile
        // layout); when it has reached its capacity (buffer_cap), and when the item closes, the buffer
        // gets its bounding circles and is published to the evaluate warps.
        const int fw = warp - 1, ftid = fw * 32 + lane;
        uint32_t it = 0, bk = 0;             // stages consumed, buffers published
        int cur = -1, count = 0, par = 0, kbuf = 0, surv = 0;
        uint32_t n_open = 0;                 // items opened
        bool have_slot = false;
        Tile<T> blk;                         // bounding circle of the block's targets
        long long fl_wait_full = 0, fl_wait_free = 0, fl_bufs = 0, fl_stages = 0;
        PROF_T0(fl_t0);
        auto write_entry = [&](unsigned char* sb, int kidx, P x, P y, T c, T s) {
            unsigned char* tb = sb + (size_t)(kidx >> 6) * kTileB;
            const int r = kidx & 63, l = r & 31, h = r >> 5;
            P* pa = reinterpret_cast<P*>(reinterpret_cast<SrcA<T>*>(tb) + l);
            T* pb = reinterpret_cast<T*>(reinterpret_cast<SrcB<T>*>(tb + 32 * sizeof(SrcA<T>)) + l);
            pa[h] = x;
            pa[2 + h] = y;
            pb[h] = c;
            pb[2 + h] = s;
        };
        // retire buffer b: wait until every evaluate warp has left it.  If it was the last buffer of an
        // item, every buffer of that item has been evaluated by now (buffer b - 1 was retired before its
        // slot was refilled): the item's sums go out, target q by filter thread q, and its accumulators
        // are cleared for the item after the next one.
        int pend_item0 = -1, pend_item1 = -1, pend_par0 = 0, pend_par1 = 0;     // per slot
        auto retire = [&](uint32_t b) {
            const int slot = (int)(b & 1u);
            PROF_T0(w0);
            if (lane == 0) mbar_wait_backoff(&freeb[slot], (b >> 1) & 1, CSF_TILED_HINT_NS);
            __syncwarp();
            PROF_ADD(fl_wait_free, w0);
            const int p_item = slot ? pend_item1 : pend_item0, p_par = slot ? pend_par1 : pend_par0;
            if (p_item >= 0) {
                const int tb = p_item % n_tblocks, cg = p_item / n_tblocks;
                T* a0 = bacc + (size_t)p_par * 2 * kBT * 2;
                for (int q = ftid; q < kBT; q += FW * 32) {
                    const long long jj = bj[p_par * kBT + q];
                    const T sx = a0[q * 2] + a0[kBT * 2 + q * 2], sy = a0[q * 2 + 1] + a0[kBT * 2 + q * 2 + 1];
                    a0[q * 2] = a0[q * 2 + 1] = a0[kBT * 2 + q * 2] = a0[kBT * 2 + q * 2 + 1] = (T)0;
                    if (jj >= 0 && (int64_t)tb * kBT + q < n_tgt) {
                        partial[((size_t)cg * n_tgt + (size_t)jj) * 2] = sx;
                        partial[((size_t)cg * n_tgt + (size_t)jj) * 2 + 1] = sy;
                    }
                }
#ifdef CSF_TILED_PROF
                if (stats && ftid == 0) stats[16 + 4 * p_item + 2] = globaltimer_ns();
#endif
                if (slot) pend_item1 = -1; else pend_item0 = -1;
            }
        };
        auto acquire = [&]() {               // the slot of buffer bk is free when buffer bk - 2 has been retired
            if (bk >= 2) retire(bk - 2);
            have_slot = true;
            count = 0;
        };
        // every append is followed by a filter_barrier before publish() is entered
        auto publish = [&](int flags) {
            const int slot = bk & 1;
            unsigned char* sb = sbuf + (size_t)slot * kDT * kTileB;
            const int n_dt = (count + kTileS - 1) / kTileS;
            {   // pad the last dynamic tile with far-away sentinels (they contribute exactly 0)
                Xycs<T> pad;
                pad_entry(pad);
                if (FIELD == 1) { pad.c = (T)0; pad.s = (T)0; }       // eccentricity 0
                for (int kidx = count + ftid; kidx < n_dt * kTileS; kidx += FW * 32)
                    write_entry(sb, kidx, pos_x(pad), pos_y(pad), pad.c, pad.s);
            }
            for (int t = fw; t < n_dt; t += FW) {            // circles of the dynamic tiles
                const int valid = min(kTileS, count - t * kTileS);
                const SrcA<T> A = reinterpret_cast<const SrcA<T>*>(sb + (size_t)t * kTileB)[lane];
                BBox<T> bb;
                if (lane < valid) bb.add(A.x0, A.y0);
                if (lane + 32 < valid) bb.add(A.x1, A.y1);
                bb.warp_reduce();
                if (lane == 0) dtile[slot * kDT + t] = bb.circle(valid);
            }
            filter_barrier<FW>();
            if (ftid == 0) {
                bdesc[slot] = make_int4(cur, n_dt, flags, par | ((kbuf & 1) << 1));
                nextq[slot] = 0;
                mbar_arrive(&ready[slot]);
            }
            if (flags & BUF_LAST) {
                if (slot) { pend_item1 = cur; pend_par1 = par; } else { pend_item0 = cur; pend_par0 = par; }
            }
            ++bk;
            ++kbuf;
            have_slot = false;
            PROF_INC(fl_bufs);
        };
        for (;;) {
            const int stage = it % kStages;
            PROF_T0(w0);
            if (lane == 0) mbar_wait_backoff(&full[stage], (it / kStages) & 1, CSF_TILED_HINT_NS);
            __syncwarp();
            PROF_ADD(fl_wait_full, w0);
            PROF_INC(fl_stages);
            const int item = hdr[stage].x, n = hdr[stage].y;
            if (n == -2) {
                acquire();
                if (bk >= 1) retire(bk - 1);                  // every item's sums are out
                publish(BUF_EXIT);
                break;
            }
            if (item != cur) {
                // open the item.  The free slot also says that the last item but one has been retired, so
                // its set of targets may be overwritten.
                acquire();
                cur = item;
                kbuf = 0;
                surv = 0;
                par = (int)(n_open++ & 1u);
                const int tb = cur % n_tblocks;
                blk = tblocks[tb];
                for (int q = ftid; q < kBT; q += FW * 32) {
                    const int64_t t = (int64_t)tb * kBT + q;
                    long long myj = -1;
                    Xycs<T> e;
                    pad_entry(e);
                    if (t < n_tgt) {
                        myj = tgt_perm ? tgt_perm[t] : t;
                        e = tgt[myj];
                    }
                    btgt[par * kBT + q] = e;
                    bj[par * kBT + q] = myj;
                }
            }
            if (n == -1) {
                if (!have_slot) acquire();
                filter_barrier<FW>();                             // targets / last appends have landed
                publish(BUF_LAST);
                if (item_cost && ftid == 0) item_cost[cur] = (unsigned int)surv;
#ifdef CSF_TILED_PROF
                if (stats && ftid == 0) stats[16 + 4 * cur + 1] = globaltimer_ns();
#endif
            } else {
                const unsigned char* base = stage_src + (size_t)stage * kST * kTileB;
                uint2* fm = fmask + stage * kST;
                // (1) filter this warp's tiles of the stage
                constexpr int kTPW = kST / FW;
                uint32_t mb0[kTPW], mb1[kTPW];
#pragma unroll
                for (int u = 0; u < kTPW; ++u) {
                    const int t = fw + u * FW;
                    mb0[u] = mb1[u] = 0;
                    if (t < n) {                              // warp-uniform
                        const int valid = (int)srec[stage * kST + t].cnt;
                        const SrcA<T> A = reinterpret_cast<const SrcA<T>*>(base + (size_t)t * kTileB)[lane];
                        const SrcB<T> B = reinterpret_cast<const SrcB<T>*>(base + (size_t)t * kTileB + 32 * sizeof(SrcA<T>))[lane];
                        bool p0, p1;
                        if (FIELD == 1) {
                            p0 = bike_reaches<T, P>(blk, A.x0, A.y0, B.c0, B.s0, lobe[0], k.tiny);
                            p1 = bike_reaches<T, P>(blk, A.x1, A.y1, B.c1, B.s1, lobe[0], k.tiny);
                        } else {
                            lobe_reaches2(blk, A, B, lobe, k.tiny, p0, p1);
                        }
                        mb0[u] = __ballot_sync(0xffffffffu, p0 && (lane < valid));
                        mb1[u] = __ballot_sync(0xffffffffu, p1 && (lane + 32 < valid));
                        if (FW > 1 && lane == 0) fm[t] = make_uint2(mb0[u], mb1[u]);
                    }
                }
                // (2) append in stream order: prefix over the stage's tiles (every filter warp computes it)
                uint2 mm = make_uint2(0u, 0u);
                if (FW > 1) {
                    filter_barrier<FW>();
                    if (lane < n) mm = fm[lane];
                } else {
#pragma unroll
                    for (int u = 0; u < kTPW; ++u)
                        if (lane == u) mm = make_uint2(mb0[u], mb1[u]);
                }
                const int c8 = __popc(mm.x) + __popc(mm.y);
                int incl = c8;
#pragma unroll
                for (int o = 1; o < kST; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int tot = __shfl_sync(0xffffffffu, incl, kST - 1);
                if (have_slot && count + tot > kDS) publish(0);     // (warp-uniform, the same in every filter warp)
                if (!have_slot) acquire();
                unsigned char* sb = sbuf + (size_t)(bk & 1) * kDT * kTileB;
                const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
                for (int u = 0; u < kTPW; ++u) {
                    const int t = fw + u * FW;
                    if (t < n && (mb0[u] | mb1[u])) {
                        const int off = count + __shfl_sync(0xffffffffu, incl - c8, t);
                        const SrcA<T> A = reinterpret_cast<const SrcA<T>*>(base + (size_t)t * kTileB)[lane];
                        const SrcB<T> B = reinterpret_cast<const SrcB<T>*>(base + (size_t)t * kTileB + 32 * sizeof(SrcA<T>))[lane];
                        if ((mb0[u] >> lane) & 1u)
                            write_entry(sb, off + __popc(mb0[u] & lt), A.x0, A.y0, B.c0, B.s0);
                        if ((mb1[u] >> lane) & 1u)
                            write_entry(sb, off + __popc(mb0[u]) + __popc(mb1[u] & lt), A.x1, A.y1, B.c1, B.s1);
                    }
                }
                count += tot;
                surv += tot;
                filter_barrier<FW>();                             // the appends have landed, the stage has been read
                if (count >= buffer_cap<kDT, (EW == kWideEW)>(kbuf)) publish(0);
            }
            if (lane == 0) mbar_arrive(&empty[stage]);
            ++it;
        }
#ifdef CSF_TILED_PROF
        if (stats && lane == 0) {
            atomicAdd(stats + 3, (unsigned long long)(clock64() - fl_t0));
            atomicAdd(stats + 4, (unsigned long long)fl_wait_full);
            atomicAdd(stats + 5, (unsigned long long)fl_wait_free);
            if (fw == 0) {
                atomicAdd(stats + 8, (unsigned long long)fl_bufs);
                atomicAdd(stats + 9, (unsigned long long)fl_stages);
            }
        }
#endif
        (void)fl_wait_full; (void)fl_wait_free; (void)fl_bufs; (void)fl_stages;
        return;
    }

    // ===== evaluate warps =====
    // Per survivor buffer the warps take the block's targets from a shared counter (the targets' headings
    // make the work per target very uneven): cull -- the target's view cone against the buffer's 32 circles,
    // one ballot (f64 build: 16-tile buffers, two targets per ballot in lanes 0-15 / 16-31); evaluate -- the surviving tiles
    // two at a time (four independent pair evaluations per lane in flight), one butterfly reduces both
    // force components (lanes 0-15: x, 16-31: y), one lane each adds them to the target's accumulator.
    // A (buffer, target) sum is formed by one warp in a fixed order whoever takes it, and buffers that
    // may be in flight together use different accumulators (parity of the buffer's number within its
    // item, parity of the item): every sum is deterministic.
    unsigned long long n_eval = 0;
    long long ev_wait = 0, ev_units = 0;
    PROF_T0(ev_t0);
    for (uint32_t bk = 0;; ++bk) {
        const int slot = (int)(bk & 1u);
        PROF_T0(w0);
        if (lane == 0) mbar_wait_backoff(&ready[slot], (bk >> 1) & 1, CSF_TILED_HINT_NS);
        __syncwarp();
        PROF_ADD(ev_wait, w0);
        const int4 d = bdesc[slot];
        if (d.z & BUF_EXIT) break;
        const int n_dt = d.y, par = d.w & 1, kpar = (d.w >> 1) & 1;
        const int tb = d.x % n_tblocks;
        const int64_t left = n_tgt - (int64_t)tb * kBT;
        const int nvalid = (int)(left < kBT ? left : kBT);
        if (n_dt > 0) {
            const unsigned char* base = sbuf + (size_t)slot * kDT * kTileB;
            const Xycs<T>* mytgt = btgt + par * kBT;
            T* acc = bacc + (size_t)(par * 2 + kpar) * kBT * 2;
            constexpr int kTPB = 32 / kDT;                   // targets per cull ballot
            const int tl = lane & (kDT - 1), half = lane / kDT;
            Tile<T> mytile;
            if (tl < n_dt) mytile = dtile[slot * kDT + tl];
            for (;;) {
                int q0 = 0;
                if (lane == 0) q0 = atomicAdd(&nextq[slot], kTPB);
                q0 = __shfl_sync(0xffffffffu, q0, 0);
                if (q0 >= nvalid) break;
                uint32_t m2;
                {
                    const int qh = q0 + half;
                    const Xycs<T> te = mytgt[qh < nvalid ? qh : q0];
                    const Tgt<T> tg = *reinterpret_cast<const Tgt<T>*>(&te);
                    const bool v = (qh < nvalid) && (tl < n_dt) && tile_visible<T, P2R>(mytile, tg, cc);
                    m2 = __ballot_sync(0xffffffffu, v);
                }
#pragma unroll 1
                for (int h = 0; h < kTPB; ++h) {
                    uint32_t mask = kTPB == 1 ? m2 : (h ? (m2 >> 16) : (m2 & 0xffffu));
                    if (mask == 0) continue;
                    PROF_INC(ev_units);
                    const int q = q0 + h;
                    if (stats) n_eval += (unsigned long long)__popc(mask) * kTileS;
                    const Xycs<T> te = mytgt[q];
                    const Tgt<T> tg = *reinterpret_cast<const Tgt<T>*>(&te);
                    typename TileAccSel<T, P2R, FIELD>::type acc0, acc1;
                    // two surviving tiles per iteration: four independent pair evaluations in flight
                    while (mask & (mask - 1)) {
                        const int t0 = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int t1 = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const unsigned char* p0 = base + (size_t)t0 * kTileB;
                        const unsigned char* p1 = base + (size_t)t1 * kTileB;
                        const SrcA<T> A0 = reinterpret_cast<const SrcA<T>*>(p0)[lane];
                        const SrcB<T> B0 = reinterpret_cast<const SrcB<T>*>(p0 + 32 * sizeof(SrcA<T>))[lane];
                        const SrcA<T> A1 = reinterpret_cast<const SrcA<T>*>(p1)[lane];
                        const SrcB<T> B1 = reinterpret_cast<const SrcB<T>*>(p1 + 32 * sizeof(SrcA<T>))[lane];
                        acc0.eval(A0, B0, tg, k);
                        acc1.eval(A1, B1, tg, k);
                    }
                    if (mask) {
                        const int t0 = __ffs(mask) - 1;
                        const unsigned char* p0 = base + (size_t)t0 * kTileB;
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
                    if ((lane & 15) == 0) acc[q * 2 + (lane >> 4)] += mine;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&freeb[slot]);
            if (d.z & BUF_LAST) atomicAdd(last_seen, 1);
        }
    }
    if (stats && n_eval && lane == 0) atomicAdd(stats, n_eval);
#ifdef CSF_TILED_PROF
    if (stats && lane == 0) {
        atomicAdd(stats + 1, (unsigned long long)(clock64() - ev_t0));
        atomicAdd(stats + 2, (unsigned long long)ev_wait);
        atomicAdd(stats + 10, (unsigned long long)ev_units);
    }
#endif
    (void)ev_wait; (void)ev_units;
}

// Items in the order of decreasing cost (one CTA, bitonic sort in shared memory): handed out in this
// order by the atomic counter, the heaviest items start first and the last ones to start are the
// lightest -- the tail of the launch (SMs idle while the last items finish) shrinks from about one
// mean item to about one of the cheapest.  cost = survivors the item's filter produced in the previous
// launch (the crowd moves a few centimetres per step).  Any order gives the same forces.
constexpr int kMaxOrderItems = 4096;
__global__ void __launch_bounds__(1024) item_order_kernel(const unsigned int* __restrict__ cost, int n,
                                                          unsigned int* __restrict__ order) {
    __shared__ unsigned long long key[kMaxOrderItems];
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x)      // descending cost, ties by index: ~cost in the high word
        key[i] = i < n ? (((unsigned long long)(~cost[i])) << 32) | (unsigned int)i : ~0ull;
    __syncthreads();
    for (int k2 = 2; k2 <= m; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = key[i], b = key[l];
                    const bool up = (i & k2) == 0;
                    if ((a > b) == up) { key[i] = b; key[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < n; i += blockDim.x) order[i] = (unsigned int)(key[i] & 0xffffffffu);
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
    constexpr int kDT = DynTiles<T>::n;
    return (size_t)kStages * kST * TileBytes<T>::v + (size_t)2 * kDT * TileBytes<T>::v +
           (size_t)(kStages * kST + 2 * kDT) * sizeof(Tile<T>) + (size_t)2 * kBT * sizeof(Xycs<T>) +
           (size_t)2 * kBT * sizeof(long long) + (size_t)(2 * kStages + 4) * sizeof(uint64_t) +
           (size_t)(kStages + 2) * sizeof(int4) + (size_t)kStages * kST * sizeof(uint2) +
           (size_t)kLobeBins * sizeof(T) + (size_t)2 * 2 * kBT * 2 * sizeof(T) + 8 * sizeof(int);
}

// Resident CTAs per SM of the tiled kernel and the SM count, per device (the shared-memory attribute
// is a per-device property of the function too).
constexpr int kMaxDevices = 64;
int g_tiled_ctas[2][2][kMaxDevices];      // [f32 / f64][narrow / wide][device]
int g_tiled_sms[kMaxDevices];
template <typename T, int FW, int EW> int tiled_shape_ctas() {
    int best = 1 << 30;
    const size_t smem = tiled_smem_bytes<T>();
    {
        auto kern = pair_tiled_kernel<T, false, FW, EW>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, (1 + FW + EW) * 32, smem);
        best = nb < best ? nb : best;
    }
    {
        auto kern = pair_tiled_kernel<T, true, FW, EW>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, (1 + FW + EW) * 32, smem);
        best = nb < best ? nb : best;
    }
    // the v0.1 Bicycle-field instances share the plan (grid sized for the TwoD instances: CTAs that are not
    // resident at once simply start later -- items come from a counter, no CTA waits for another)
    cudaFuncSetAttribute(pair_tiled_kernel<T, false, FW, EW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(pair_tiled_kernel<T, true, FW, EW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return best < 1 ? 1 : best;
}
// resident CTAs per SM of the narrow (wide = false) or wide CTA shape on the current device
template <typename T> int tiled_ctas(bool wide, int* sm_count) {
    const int idx = sizeof(T) == 4 ? 0 : 1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    if (g_tiled_ctas[idx][0][dev] == 0) {
        g_tiled_ctas[idx][0][dev] = tiled_shape_ctas<T, kNarrowFW, kNarrowEW>();
        g_tiled_ctas[idx][1][dev] = tiled_shape_ctas<T, kWideFW, kWideEW>();
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        g_tiled_sms[dev] = sms > 0 ? sms : 148;
    }
    if (sm_count) *sm_count = g_tiled_sms[dev];
    return g_tiled_ctas[idx][wide ? 1 : 0][dev];
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Work decomposition: target blocks of kBT targets; if there are too few blocks to keep every CTA slot
// busy, the list of chunks near a block is dealt out to several items, and below a few blocks per SM the
// wide CTA shape is used (see the top of the file).
struct TiledPlan { int n_tblocks, n_groups, grid; bool wide; int64_t n_tiles, n_chunks; };
template <typename T> TiledPlan tiled_plan(int64_t n_src, int64_t n_tgt) {
    static const int env_groups = env_int("CSF_TILED_GROUPS", 0), env_ipw = env_int("CSF_TILED_ITEMS_PER_SLOT", 3),
                     env_max = env_int("CSF_TILED_MAX_GROUPS", 16), env_wide = env_int("CSF_TILED_WIDE", -1),
                     env_wide_below = env_int("CSF_TILED_WIDE_BELOW", 3);
    TiledPlan p;
    p.n_tiles = (n_src + kTileS - 1) / kTileS;
    p.n_chunks = (p.n_tiles + kCT - 1) / kCT;
    const int64_t tb = (n_tgt + kBT - 1) / kBT;
    int sms = 0;
    tiled_ctas<T>(false, &sms);
    // fewer than env_wide_below target blocks per SM: item latency, not throughput, bounds the launch
    p.wide = sizeof(T) == 4 && (env_wide >= 0 ? env_wide != 0 : tb < (int64_t)env_wide_below * sms);
    const int64_t slots = (int64_t)tiled_ctas<T>(p.wide, &sms) * sms;
    // 64 targets per block balance the lobe filter's cost (per block) against its selectivity (block
    // radius).  Narrow shape: items must be plentiful -- a few per CTA slot -- because their cost follows
    // the local density of the crowd and an item is the unit of dynamic scheduling; but every split of a
    // block's chunk list halves the survivors per item, hence the tiles per (target, buffer) unit, whose
    // fixed cost (cull, reductions) is a fifth of the evaluate warps' instructions.  Measured at N = 65,536
    // with 296 CTA slots (K1, us): 1,024 items 291, 2,048 items 308, 3,072 items 318 -- three items per slot.  Wide shape: one item
    // per SM at a time; the chunks near a block are dealt out to several items only when there are fewer
    // blocks than SMs.  (Each chunk is still filtered once per block either way.)
    const int64_t want = p.wide ? slots : (int64_t)env_ipw * slots;
    int64_t groups = (want + tb - 1) / tb;
    if (groups > env_max) groups = env_max;
    if (env_groups >= 1) groups = env_groups;
    if (groups < 1) groups = 1;
    if (groups > kTMaxGroups) groups = kTMaxGroups;
    if (groups > p.n_chunks) groups = p.n_chunks;
    p.n_tblocks = (int)tb;
    p.n_groups = (int)groups;
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

// v0.1 Bicycle field (vehicle.py:1054-1147): |F| = (p_0 / p_decay) exp(-b) sqrt(g^2 + (e sin(phi0))^2 / (1 - e^2)),
// g = (1 - e cos(phi0)) / sqrt(1 - e^2), b = rho g / p_decay, and e <= 0.7 (:1062-1064): the square root is at
// most 2.58 < 2^1.5, so |F| < 2^-cutoff_log2 p_0 / p_decay wherever b > (cutoff_log2 + 1.5) ln 2, i.e. beyond
// rho = K sqrt(1 - e^2) / (1 - e cos(phi0)) with K [m] below; straight ahead at e = 0.7 that is K sqrt(1.7 / 0.3).
constexpr double kBikeEmax = 0.7;
const double kBikeAhead = sqrt((1.0 + kBikeEmax) / (1.0 - kBikeEmax));
double bike_reach_K(const CsfFieldParams* fp) {
    const double bits = fp->cutoff_log2 > 0.0 ? fp->cutoff_log2 : 40.0;
    return (bits + 1.5) * 0.6931471805599453 * fp->p_decay;
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
    if (fp->field_kind == 1) {
        // v0.1 Bicycle field: lobe[0] = K of bike_reaches, dmax = the reach straight ahead at the largest eccentricity
        if (is_f32) {
            const double K = bike_reach_K(fp) * 1.0001 / fp->q_scale;
            if (K * kBikeAhead < 3.0e9) {
                c.lobe[0] = (T)K;
                c.dmax = (T)(K * kBikeAhead);
            }
        }
        return c;
    }
    if (is_f32) {
        // exp(-d * rate) = 2^-cutoff_log2  ->  reach d.  The bounds cost ~1 ms of host time per parameter
        // set: a small cache (several source classes alternate within a step)
        constexpr int NC = 512, kCache = 8;
        struct Entry { CsfFieldParams fp; double rate; double rate_c[NC]; bool ok; bool used; };
        static Entry cache[kCache];
        static int next_slot = 0;
        static std::mutex cache_mutex;                   // engines on several host threads share the cache
        std::lock_guard<std::mutex> lock(cache_mutex);
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
int tile_sources(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles, cudaStream_t st,
                 const void* speed = nullptr, double v_max = 1.0) {
    if (n <= 0) return 0;
    const int64_t n_tiles = (n + kTileS - 1) / kTileS, n_chunks = (n_tiles + kCT - 1) / kCT;
    tile_sources_kernel<T><<<(unsigned)n_chunks, kCT * 32, 0, st>>>(
        (const Xycs<T>*)xycs, n, perm, (unsigned char*)sorted, (Tile<T>*)tiles, n_tiles, (const T*)speed, (T)v_max);
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
               const unsigned int* item_order, unsigned int* item_cost, unsigned long long* stats, int flags,
               cudaStream_t st) {
    if (n_tgt <= 0) return 0;
    const double amplitude = field_amplitude(fp);
    if (n_src <= 0 || amplitude == 0.0) {
        if (flags & CSF_TILED_NO_REDUCE) {
            csf_set_error("csf_pair_forces_tiled: CSF_TILED_NO_REDUCE needs sources and f_0 != 0", cudaErrorInvalidValue);
            return -(int)cudaErrorInvalidValue;
        }
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
    if (!(flags & CSF_TILED_PREPARED)) {
        block_bounds_kernel<T><<<(unsigned)(((int64_t)pl.n_tblocks * 32 + 127) / 128), 128, 0, st>>>(
            (const Xycs<T>*)tgt, tgt_perm, n_tgt, kBT, tblocks, pl.n_tblocks, counter);
        CSF_CHECK_LAUNCH("block_bounds_kernel");
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)pl.grid);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (flags & CSF_TILED_PDL) ? 1 : 0;
    const unsigned char* sorted_c = (const unsigned char*)sorted;
    const Tile<T>* tiles_c = (const Tile<T>*)tiles;
    const Xycs<T>* tgt_c = (const Xycs<T>*)tgt;
    const Tile<T>* tblocks_c = tblocks;
#define CSF_TILED_LAUNCH(P2R_, FW_, EW_, FIELD_)                                                                     \
    cfg.blockDim = dim3((1 + FW_ + EW_) * 32);                                                                       \
    cudaLaunchKernelEx(&cfg, pair_tiled_kernel<T, P2R_, FW_, EW_, FIELD_>, sorted_c, tiles_c, pl.n_tiles, tgt_c, tgt_perm, \
                       n_tgt, tblocks_c, k, cc, partial, pl.n_groups, pl.n_tblocks, counter, item_order, item_cost, stats)
#define CSF_TILED_LAUNCH_SHAPE(P2R_, FIELD_)                                                                         \
    if (pl.wide) { CSF_TILED_LAUNCH(P2R_, kWideFW, kWideEW, FIELD_); }                                               \
    else { CSF_TILED_LAUNCH(P2R_, kNarrowFW, kNarrowEW, FIELD_); }
    if (fp->field_kind == 1) {
        if (fp->p2r) { CSF_TILED_LAUNCH_SHAPE(true, 1) } else { CSF_TILED_LAUNCH_SHAPE(false, 1) }
    } else {
        if (fp->p2r) { CSF_TILED_LAUNCH_SHAPE(true, 0) } else { CSF_TILED_LAUNCH_SHAPE(false, 0) }
    }
#undef CSF_TILED_LAUNCH_SHAPE
#undef CSF_TILED_LAUNCH
    CSF_CHECK_LAUNCH("pair_tiled_kernel");
    if (flags & CSF_TILED_NO_REDUCE) return 0;          // the caller sums partial[group][target][2] * f_0 itself
    const int64_t n2 = n_tgt * 2;
    reduce_groups_kernel<T><<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(partial, pl.n_groups, n_tgt, (T)amplitude, frep,
                                                                           accumulate);
    CSF_CHECK_LAUNCH("reduce_groups_kernel");
    return 0;
}

template <typename T>
int tiled_prepare(const void* xycs, int64_t n_src, const int64_t* perm, void* sorted, void* tiles, const void* tgt,
                  const int64_t* tgt_perm, int64_t n_tgt, void* ws, size_t wsb, const CsfPeerComm* comm, cudaStream_t st) {
    if (n_src <= 0 || n_tgt <= 0) return 0;
    const TiledPlan pl = tiled_plan<T>(n_src, n_tgt);
    if (ws == nullptr || wsb < kWsHeader + tiled_ws_blocks_bytes<T>(pl)) {
        csf_set_error("csf_tiled_prepare: workspace too small", cudaErrorInvalidValue);
        return -(int)cudaErrorInvalidValue;
    }
    CsfPeerComm c;
    memset(&c, 0, sizeof(c));
    if (comm) c = *comm;
    const int64_t bb_ctas = ((int64_t)pl.n_tblocks + kCT - 1) / kCT;
    tiled_prepare_kernel<T><<<(unsigned)(pl.n_chunks + bb_ctas), kCT * 32, 0, st>>>(
        (const Xycs<T>*)xycs, n_src, perm, (unsigned char*)sorted, (Tile<T>*)tiles, pl.n_tiles, pl.n_chunks,
        (const Xycs<T>*)tgt, tgt_perm, n_tgt, reinterpret_cast<Tile<T>*>((unsigned char*)ws + kWsHeader), pl.n_tblocks,
        reinterpret_cast<unsigned int*>(ws), c);
    CSF_CHECK_LAUNCH("tiled_prepare_kernel");
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
    if (fp->field_kind == 1) return bike_reach_K(fp) * kBikeAhead;
    const double rate = field_min_decay_rate(fp);
    const double bits = fp->cutoff_log2 > 0.0 ? fp->cutoff_log2 : 40.0;
    return rate > 0.0 ? bits * 0.6931471805599453 / rate : INFINITY;
}
int csf_field_reach_table(const CsfFieldParams* fp, int n_bins, double* reach_m) {
    // host-only: the reach table of the lobe filter in metres (bin b covers cos(phi) in
    // [-1 + 2b/n_bins, -1 + 2(b+1)/n_bins]); n_bins must be kLobeBins
    if (n_bins != kLobeBins) return -1;
    if (fp->field_kind == 1) {      // the ellipse of bike_reaches at the largest eccentricity, at each bin's upper edge
        const double K = bike_reach_K(fp);
        for (int b = 0; b < kLobeBins; ++b)
            reach_m[b] = K * sqrt(1.0 - kBikeEmax * kBikeEmax) / (1.0 - kBikeEmax * (-1.0 + 2.0 * (b + 1) / kLobeBins));
        return 0;
    }
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
int csf_tile_sources_bicycle_f32(const void* xycs, const float* speed, double v_max, int64_t n, const int64_t* perm,
                                 void* sorted, void* tiles, csf_stream_t st) {
    return tile_sources<float>(xycs, n, perm, sorted, tiles, (cudaStream_t)st, speed, v_max);
}
int csf_tile_sources_bicycle_f64(const void* xycs, const double* speed, double v_max, int64_t n, const int64_t* perm,
                                 void* sorted, void* tiles, csf_stream_t st) {
    return tile_sources<double>(xycs, n, perm, sorted, tiles, (cudaStream_t)st, speed, v_max);
}
int csf_pair_forces_tiled_f32(const void* sorted, const void* tiles, int64_t n_src, const void* tgt,
                              const int64_t* tgt_perm, int64_t n_tgt, const CsfFieldParams* fp, float* frep,
                              int accumulate, void* ws, size_t wsb, const unsigned int* item_order,
                              unsigned int* item_cost, unsigned long long* stats, int flags, csf_stream_t st) {
    return pair_tiled<float>(sorted, tiles, n_src, tgt, tgt_perm, n_tgt, fp, frep, accumulate, ws, wsb, item_order,
                             item_cost, stats, flags, (cudaStream_t)st);
}
int csf_pair_forces_tiled_f64(const void* sorted, const void* tiles, int64_t n_src, const void* tgt,
                              const int64_t* tgt_perm, int64_t n_tgt, const CsfFieldParams* fp, double* frep,
                              int accumulate, void* ws, size_t wsb, const unsigned int* item_order,
                              unsigned int* item_cost, unsigned long long* stats, int flags, csf_stream_t st) {
    return pair_tiled<double>(sorted, tiles, n_src, tgt, tgt_perm, n_tgt, fp, frep, accumulate, ws, wsb, item_order,
                              item_cost, stats, flags, (cudaStream_t)st);
}
int csf_tiled_prepare_f32(const void* xycs, int64_t n_src, const int64_t* perm, void* sorted, void* tiles, const void* tgt,
                          const int64_t* tgt_perm, int64_t n_tgt, void* ws, size_t wsb, const CsfPeerComm* comm,
                          csf_stream_t st) {
    return tiled_prepare<float>(xycs, n_src, perm, sorted, tiles, tgt, tgt_perm, n_tgt, ws, wsb, comm, (cudaStream_t)st);
}
int csf_tiled_prepare_f64(const void* xycs, int64_t n_src, const int64_t* perm, void* sorted, void* tiles, const void* tgt,
                          const int64_t* tgt_perm, int64_t n_tgt, void* ws, size_t wsb, const CsfPeerComm* comm,
                          csf_stream_t st) {
    return tiled_prepare<double>(xycs, n_src, perm, sorted, tiles, tgt, tgt_perm, n_tgt, ws, wsb, comm, (cudaStream_t)st);
}
int csf_tiled_num_groups(int64_t n_src, int64_t n_tgt, int elem_bytes) {
    if (n_src <= 0 || n_tgt <= 0) return 0;
    return elem_bytes == 4 ? tiled_plan<float>(n_src, n_tgt).n_groups : tiled_plan<double>(n_src, n_tgt).n_groups;
}
size_t csf_tiled_partial_offset(int64_t n_src, int64_t n_tgt, int elem_bytes) {
    if (n_src <= 0 || n_tgt <= 0) return 0;
    if (elem_bytes == 4) return kWsHeader + tiled_ws_blocks_bytes<float>(tiled_plan<float>(n_src, n_tgt));
    return kWsHeader + tiled_ws_blocks_bytes<double>(tiled_plan<double>(n_src, n_tgt));
}
int64_t csf_tiled_num_items(int64_t n_src, int64_t n_tgt, int elem_bytes) {
    if (n_src <= 0 || n_tgt <= 0) return 0;
    const TiledPlan pl = elem_bytes == 4 ? tiled_plan<float>(n_src, n_tgt) : tiled_plan<double>(n_src, n_tgt);
    return (int64_t)pl.n_tblocks * pl.n_groups;
}
int csf_tiled_item_order(const unsigned int* item_cost, int64_t n_items, unsigned int* item_order, csf_stream_t st) {
    if (n_items <= 0) return 0;
    if (n_items > kMaxOrderItems) {
        csf_set_error("csf_tiled_item_order: more than 4096 items (pass item_order = NULL instead)", cudaErrorInvalidValue);
        return -(int)cudaErrorInvalidValue;
    }
    item_order_kernel<<<1, 1024, 0, (cudaStream_t)st>>>(item_cost, (int)n_items, item_order);
    CSF_CHECK_LAUNCH("item_order_kernel");
    return 0;
}

}  // extern "C"
