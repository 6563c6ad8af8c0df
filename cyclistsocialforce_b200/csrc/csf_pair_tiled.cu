// K1 (tiled): all-pairs repulsive force with exact field-of-view culling.
//
// ~2/3 of the ordered pairs of a crowd are masked by the target's field of view
// (reference intersection.py:733-736; hfov = 2 pi / 3).  The mask depends on the target's
// heading, so a thread-per-target kernel cannot skip them (every lane sees a different cone).
// This kernel turns the mapping around:
//   * sources are kept in a spatially sorted copy (Morton order, refreshed by the host every
//     few steps) cut into tiles of 64 with a bounding circle each (csf_tile_sources_*);
//   * a warp works on ONE target at a time: the lanes first test one tile each against the
//     target's view cone expanded by the tile radius (a conservative, warp-uniform decision
//     after a ballot), then the lanes evaluate the surviving tiles two sources per lane with
//     the same pair_eval as the dense kernel (which still applies the exact per-pair mask);
//   * lane partial sums are combined with warp shuffles once per (target, source chunk).
// Culled pairs are exactly pairs whose mask is 0, so the result equals the dense kernel's up
// to the order of summation.  In the f32 build tiles whose nearest point is so far that every
// contribution is below 2^-40 f_0 are skipped too (bounded, documented truncation; never in f64).
//
// Source chunks (16 tiles + their records) are streamed into a shared-memory ring by a
// producer warp with TMA bulk copies, exactly like the dense kernel.
#include "csf_common.cuh"
#include "csf_pair_common.cuh"

namespace {

#ifndef CSF_TILED_WARPS
#define CSF_TILED_WARPS 8        // consumer warps per CTA
#endif
#ifndef CSF_TILED_TPW
#define CSF_TILED_TPW 16         // targets per warp per work item
#endif
#ifndef CSF_TILED_CHUNK_TILES
#define CSF_TILED_CHUNK_TILES 16 // tiles per shared-memory stage (<= 32: one tile per lane in the cull test)
#endif
#ifndef CSF_TILED_MINB
#define CSF_TILED_MINB 2
#endif
#ifndef CSF_TILED_ILP4
#define CSF_TILED_ILP4 1         // evaluate two surviving tiles per loop iteration
#endif
constexpr int kTW = CSF_TILED_WARPS;
constexpr int kTThreads = (kTW + 1) * 32;
constexpr int kTPW = CSF_TILED_TPW;
constexpr int kTB = kTW * kTPW;          // targets per work item
constexpr int kTileS = 64;               // sources per tile (2 per lane)
constexpr int kCT = CSF_TILED_CHUNK_TILES;
constexpr int kCS = kCT * kTileS;        // sources per chunk
constexpr int kTStages = 3;
constexpr int kTMaxGroups = 64;
static_assert(kTPW * 2 <= 32, "per-warp accumulator update uses one lane per scalar");
static_assert(kCT <= 32, "one tile per lane in the cull test");

template <typename T> struct Tile;
template <> struct __align__(16) Tile<float> { int32_t cx, cy; float R; int32_t cnt; };
template <> struct __align__(16) Tile<double> { double cx, cy, R; int64_t cnt; };

template <typename T> struct CullConst {
    T ca, sa;    // cos / sin of hfov/2
    T dmax;      // distance (payload units) beyond which a whole tile contributes < 2^-40 f_0 (inf: never)
};

// ---- tile builder: one warp per tile -----------------------------------------------------------
__device__ __forceinline__ void pad_entry(Xycs<float>& e) { e.xq = 1 << 30; e.yq = 1 << 30; e.c = 1.f; e.s = 0.f; }
__device__ __forceinline__ void pad_entry(Xycs<double>& e) { e.x = 1e150; e.y = 1e150; e.c = 1.0; e.s = 0.0; }
__device__ __forceinline__ void tile_bounds(const Xycs<float>& a, const Xycs<float>& b, bool va, bool vb, int cnt,
                                            Tile<float>* out, int lane) {
    int xmin = va ? a.xq : INT32_MAX, xmax = va ? a.xq : INT32_MIN, ymin = va ? a.yq : INT32_MAX,
        ymax = va ? a.yq : INT32_MIN;
    if (vb) { xmin = min(xmin, b.xq); xmax = max(xmax, b.xq); ymin = min(ymin, b.yq); ymax = max(ymax, b.yq); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (lane == 0) {
        Tile<float> t;
        t.cx = (int)(((int64_t)xmin + xmax) >> 1);
        t.cy = (int)(((int64_t)ymin + ymax) >> 1);
        const float hx = (float)((int64_t)xmax - xmin) * 0.5f + 1.f, hy = (float)((int64_t)ymax - ymin) * 0.5f + 1.f;
        t.R = sqrtf(hx * hx + hy * hy) * 1.000001f + 1.f;
        t.cnt = cnt;
        *out = t;
    }
}
__device__ __forceinline__ void tile_bounds(const Xycs<double>& a, const Xycs<double>& b, bool va, bool vb, int cnt,
                                            Tile<double>* out, int lane) {
    double xmin = va ? a.x : 1e300, xmax = va ? a.x : -1e300, ymin = va ? a.y : 1e300, ymax = va ? a.y : -1e300;
    if (vb) { xmin = fmin(xmin, b.x); xmax = fmax(xmax, b.x); ymin = fmin(ymin, b.y); ymax = fmax(ymax, b.y); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (lane == 0) {
        Tile<double> t;
        t.cx = 0.5 * (xmin + xmax);
        t.cy = 0.5 * (ymin + ymax);
        const double hx = 0.5 * (xmax - xmin), hy = 0.5 * (ymax - ymin);
        t.R = sqrt(hx * hx + hy * hy) * (1.0 + 1e-12) + 1e-9;
        t.cnt = cnt;
        *out = t;
    }
}

// sorted[t*64 + l] = xycs[perm[t*64 + l]] (perm == nullptr: identity); entries past n are padded
// with a far-away sentinel that contributes exactly 0.
template <typename T>
__global__ void tile_sources_kernel(const Xycs<T>* __restrict__ xycs, int64_t n, const int64_t* __restrict__ perm,
                                    Xycs<T>* __restrict__ sorted, Tile<T>* __restrict__ tiles, int64_t n_tiles) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    const int64_t i0 = t * kTileS + lane, i1 = i0 + 32;
    Xycs<T> a, b;
    const bool va = i0 < n, vb = i1 < n;
    if (va) a = xycs[perm ? perm[i0] : i0]; else pad_entry(a);
    if (vb) b = xycs[perm ? perm[i1] : i1]; else pad_entry(b);
    sorted[i0] = a;
    sorted[i1] = b;
    const int64_t rem = n - t * kTileS;
    tile_bounds(a, b, va, vb, (int)(rem < kTileS ? rem : kTileS), tiles + t, lane);
}

// Morton key of a payload position (host sorts the keys; any stable order works)
__device__ __forceinline__ uint32_t part1by1(uint32_t x) {
    x &= 0x0000ffff;
    x = (x | (x << 8)) & 0x00ff00ff;
    x = (x | (x << 4)) & 0x0f0f0f0f;
    x = (x | (x << 2)) & 0x33333333;
    x = (x | (x << 1)) & 0x55555555;
    return x;
}
__device__ __forceinline__ void key_xy(const Xycs<float>& e, double, double, double, uint32_t& kx, uint32_t& ky) {
    kx = ((uint32_t)(e.xq + (1 << 30))) >> 15;  // 16 bits of the 31-bit range
    ky = ((uint32_t)(e.yq + (1 << 30))) >> 15;
}
__device__ __forceinline__ void key_xy(const Xycs<double>& e, double x0, double y0, double inv_cell, uint32_t& kx,
                                       uint32_t& ky) {
    kx = (uint32_t)fmin(fmax((e.x - x0) * inv_cell, 0.0), 65535.0);
    ky = (uint32_t)fmin(fmax((e.y - y0) * inv_cell, 0.0), 65535.0);
}
template <typename T>
__global__ void morton_kernel(const Xycs<T>* __restrict__ xycs, int64_t n, double x0, double y0, double inv_cell,
                              int64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t kx, ky;
    key_xy(xycs[i], x0, y0, inv_cell, kx, ky);
    keys[i] = (int64_t)(part1by1(kx) | (part1by1(ky) << 1));
}

// ---- cull test: tile circle (c, R) against target j's view cone --------------------------------
// u = c - p_j in the target's frame: along = u . h_j, cross = u x h_j.  Every point of the circle
// is outside the (closed) cone of half angle a if  |cross| cos a - along sin a > R  (distance to the
// supporting half-plane; also valid for a >= 90 deg, where it states that the circle lies in the
// complementary cone).  p2r additionally hides sources to the left of the heading.
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

__device__ __forceinline__ Tgt<float> bcast(const Tgt<float>& v, int src) {
    Tgt<float> r;
    r.xq = __shfl_sync(0xffffffffu, v.xq, src);
    r.yq = __shfl_sync(0xffffffffu, v.yq, src);
    r.c = __shfl_sync(0xffffffffu, v.c, src);
    r.s = __shfl_sync(0xffffffffu, v.s, src);
    return r;
}
__device__ __forceinline__ Tgt<double> bcast(const Tgt<double>& v, int src) {
    Tgt<double> r;
    r.x = __shfl_sync(0xffffffffu, v.x, src);
    r.y = __shfl_sync(0xffffffffu, v.y, src);
    r.c = __shfl_sync(0xffffffffu, v.c, src);
    r.s = __shfl_sync(0xffffffffu, v.s, src);
    return r;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- the tiled pair kernel ---------------------------------------------------------------------------
// item -> (target block tb = item % n_tblocks, chunk group cg = item / n_tblocks); a group is
// `group_chunks` consecutive chunks of kCT tiles.  partial[cg][target][2].
template <typename T, bool P2R>
__global__ void __launch_bounds__(kTThreads, CSF_TILED_MINB)
pair_tiled_kernel(const Xycs<T>* __restrict__ sorted, const Tile<T>* __restrict__ tiles, int64_t n_tiles,
                  const Xycs<T>* __restrict__ tgt, int64_t n_tgt, PairConst<T> k, CullConst<T> cc,
                  T* __restrict__ partial, int group_chunks, int n_groups, int n_tblocks,
                  unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr size_t kStageBytes = (size_t)kCS * sizeof(Xycs<T>) + (size_t)kCT * sizeof(Tile<T>);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kTStages * kStageBytes);
    uint64_t* empty = full + kTStages;
    T* acc = reinterpret_cast<T*>(empty + kTStages);  // [kTW][kTPW][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t n_items = (int64_t)n_tblocks * n_groups;
    const int64_t n_chunks = (n_tiles + kCT - 1) / kCT;

    if (warp == kTW) {
        // ===== producer: stream the chunks of every item's chunk group =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int64_t cg = item / n_tblocks;
                const int64_t c_end = min(n_chunks, (cg + 1) * (int64_t)group_chunks);
                for (int64_t ch = cg * group_chunks; ch < c_end; ++ch, ++it) {
                    const int stage = it % kTStages;
                    const uint32_t phase = (it / kTStages) & 1;
                    mbar_wait(&empty[stage], phase ^ 1);
                    const int64_t t0 = ch * kCT;
                    const uint32_t nt = (uint32_t)min((int64_t)kCT, n_tiles - t0);
                    const uint32_t bsrc = nt * kTileS * (uint32_t)sizeof(Xycs<T>), btile = nt * (uint32_t)sizeof(Tile<T>);
                    unsigned char* base = smem_raw + stage * kStageBytes;
                    mbar_expect_tx(&full[stage], bsrc + btile);
                    tma_bulk_g2s(base, sorted + t0 * kTileS, bsrc, &full[stage]);
                    tma_bulk_g2s(base + (size_t)kCS * sizeof(Xycs<T>), tiles + t0, btile, &full[stage]);
                }
            }
        }
        return;
    }

    // ===== consumer warps: one target at a time =====
    uint32_t it = 0;
    unsigned long long n_eval = 0;
    T* wacc = acc + warp * kTPW * 2;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t tb = item % n_tblocks, cg = item / n_tblocks;
        const int64_t c_end = min(n_chunks, (cg + 1) * (int64_t)group_chunks);
        const int64_t t_first = tb * kTB + warp * kTPW;
        if (lane < kTPW * 2) wacc[lane] = (T)0;
        // lane q keeps target q of this warp in registers for the whole item
        Tgt<T> mine;
        {
            const int64_t jm = min(t_first + (lane % kTPW), n_tgt - 1);
            const Xycs<T> e = tgt[jm < 0 ? 0 : jm];
            mine = *reinterpret_cast<const Tgt<T>*>(&e);
        }
        __syncwarp();
        for (int64_t ch = cg * group_chunks; ch < c_end; ++ch, ++it) {
            const int stage = it % kTStages;
            const uint32_t phase = (it / kTStages) & 1;
            const int nt = (int)min((int64_t)kCT, n_tiles - ch * kCT);
            mbar_wait(&full[stage], phase);
            const Xycs<T>* src = reinterpret_cast<const Xycs<T>*>(smem_raw + stage * kStageBytes);
            const Tile<T>* trec = reinterpret_cast<const Tile<T>*>(smem_raw + stage * kStageBytes +
                                                                  (size_t)kCS * sizeof(Xycs<T>));
            Tile<T> mytile;
            if (lane < nt) mytile = trec[lane];
#pragma unroll 1
            for (int q = 0; q < kTPW; ++q) {
                const int64_t j = t_first + q;
                if (j >= n_tgt) break;
                const Tgt<T> tg = bcast(mine, q);
                const bool v = (lane < nt) && tile_visible<T, P2R>(mytile, tg, cc);
                uint32_t mask = __ballot_sync(0xffffffffu, v);
                if (mask == 0) continue;
                T ax = (T)0, ay = (T)0;
                if (stats) n_eval += (unsigned long long)__popc(mask) * kTileS;
#if CSF_TILED_ILP4
                // two surviving tiles per iteration: four independent pair evaluations in flight
                T bx = (T)0, by = (T)0;
                while (mask & (mask - 1)) {
                    const int t0 = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int t1 = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const Xycs<T> s0 = src[t0 * kTileS + lane];
                    const Xycs<T> s1 = src[t0 * kTileS + 32 + lane];
                    const Xycs<T> s2 = src[t1 * kTileS + lane];
                    const Xycs<T> s3 = src[t1 * kTileS + 32 + lane];
                    pair_eval<T, P2R>(s0, tg, k, ax, ay);
                    pair_eval<T, P2R>(s1, tg, k, bx, by);
                    pair_eval<T, P2R>(s2, tg, k, ax, ay);
                    pair_eval<T, P2R>(s3, tg, k, bx, by);
                }
                if (mask) {
                    const int t = __ffs(mask) - 1;
                    const Xycs<T> s0 = src[t * kTileS + lane];
                    const Xycs<T> s1 = src[t * kTileS + 32 + lane];
                    pair_eval<T, P2R>(s0, tg, k, ax, ay);
                    pair_eval<T, P2R>(s1, tg, k, bx, by);
                }
                ax += bx;
                ay += by;
#else
                while (mask) {
                    const int t = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const Xycs<T> s0 = src[t * kTileS + lane];
                    const Xycs<T> s1 = src[t * kTileS + 32 + lane];
                    pair_eval<T, P2R>(s0, tg, k, ax, ay);
                    pair_eval<T, P2R>(s1, tg, k, ax, ay);
                }
#endif
                ax = warp_sum(ax);
                ay = warp_sum(ay);
                if (lane == 0) {
                    wacc[q * 2] += ax;
                    wacc[q * 2 + 1] += ay;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }
        __syncwarp();
        if (lane < kTPW * 2) {
            const int64_t j = t_first + (lane >> 1);
            if (j < n_tgt) partial[((size_t)cg * n_tgt + j) * 2 + (lane & 1)] = wacc[lane];
        }
        __syncwarp();
    }
    if (stats && lane == 0 && n_eval) atomicAdd(stats, n_eval);
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
    return kTStages * ((size_t)kCS * sizeof(Xycs<T>) + (size_t)kCT * sizeof(Tile<T>)) + 2 * kTStages * sizeof(uint64_t) +
           (size_t)kTW * kTPW * 2 * sizeof(T);
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

struct TiledPlan { int n_tblocks, n_groups, group_chunks, grid; int64_t n_tiles; };
template <typename T> TiledPlan tiled_plan(int64_t n_src, int64_t n_tgt) {
    TiledPlan p;
    p.n_tiles = (n_src + kTileS - 1) / kTileS;
    const int64_t n_chunks = (p.n_tiles + kCT - 1) / kCT;
    const int64_t tb = (n_tgt + kTB - 1) / kTB;
    const int64_t slots = (int64_t)csf_sm_count() * tiled_ctas<T>();
    int64_t groups = (64 * slots + tb - 1) / tb;      // items >= 64 x slots (item cost varies with culling)
    if (groups < 1) groups = 1;
    if (groups > kTMaxGroups) groups = kTMaxGroups;
    if (groups > n_chunks) groups = n_chunks;
    const int64_t gc = (n_chunks + groups - 1) / groups;
    groups = (n_chunks + gc - 1) / gc;
    p.n_tblocks = (int)tb;
    p.n_groups = (int)groups;
    p.group_chunks = (int)gc;
    const int64_t items = tb * groups;
    p.grid = (int)(items < slots ? items : slots);
    return p;
}

template <typename T> CullConst<T> make_cull(const CsfFieldParams* fp, bool is_f32) {
    CullConst<T> c;
    const double a = fmin(fp->hfov * 0.5, CSF_PI);
    c.ca = (T)cos(a);
    c.sa = (T)sin(a);
    if (a >= CSF_PI) { c.ca = (T)-1; c.sa = (T)0; }
    if (is_f32) {
        // every pair of the tile has rho >= d - R; exponent rho q / sigma >= rho * qmin / sigma_max
        const double emax = fmax(fabs(fp->e_0), fabs(fp->e_0 - fp->e_1));
        const double qmin = sqrt(fmax(1.0 - emax * emax, 1e-12));
        const double smax = fmax(fp->sigma_0, fp->sigma_0 + fp->sigma_1);
        const double d_m = 40.0 * 0.6931471805599453 * smax / qmin;   // metres: exp(-d qmin/smax) = 2^-40
        c.dmax = (T)(d_m / fp->q_scale);
    } else {
        c.dmax = (T)1e150;
    }
    return c;
}

template <typename T>
int tile_sources(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles, cudaStream_t st) {
    if (n <= 0) return 0;
    const int64_t n_tiles = (n + kTileS - 1) / kTileS;
    const int64_t threads = n_tiles * 32;
    tile_sources_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(
        (const Xycs<T>*)xycs, n, perm, (Xycs<T>*)sorted, (Tile<T>*)tiles, n_tiles);
    CSF_CHECK_LAUNCH("tile_sources_kernel");
    return 0;
}

template <typename T>
int pair_tiled(const void* sorted, const void* tiles, int64_t n_src, const void* tgt, int64_t n_tgt,
               const CsfFieldParams* fp, T* frep, int accumulate, void* ws, size_t wsb, unsigned long long* stats,
               cudaStream_t st) {
    if (n_tgt <= 0) return 0;
    if (n_src <= 0 || fp->f_0 == 0.0) {
        if (!accumulate) cudaMemsetAsync(frep, 0, sizeof(T) * 2 * n_tgt, st);
        return 0;
    }
    const TiledPlan pl = tiled_plan<T>(n_src, n_tgt);
    const size_t need = (size_t)pl.n_groups * n_tgt * 2 * sizeof(T);
    if (ws == nullptr || wsb < need) {
        csf_set_error("csf_pair_forces_tiled: workspace too small", cudaErrorInvalidValue);
        return -(int)cudaErrorInvalidValue;
    }
    const PairConst<T> k = make_const<T>(fp, sizeof(T) == 4);
    const CullConst<T> cc = make_cull<T>(fp, sizeof(T) == 4);
    const size_t smem = tiled_smem_bytes<T>();
    T* partial = reinterpret_cast<T*>(ws);
    if (fp->p2r)
        pair_tiled_kernel<T, true><<<pl.grid, kTThreads, smem, st>>>(
            (const Xycs<T>*)sorted, (const Tile<T>*)tiles, pl.n_tiles, (const Xycs<T>*)tgt, n_tgt, k, cc, partial,
            pl.group_chunks, pl.n_groups, pl.n_tblocks, stats);
    else
        pair_tiled_kernel<T, false><<<pl.grid, kTThreads, smem, st>>>(
            (const Xycs<T>*)sorted, (const Tile<T>*)tiles, pl.n_tiles, (const Xycs<T>*)tgt, n_tgt, k, cc, partial,
            pl.group_chunks, pl.n_groups, pl.n_tblocks, stats);
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
int64_t csf_tiled_num_tiles(int64_t n_src) { return (n_src + kTileS - 1) / kTileS; }
int csf_tiled_tile_bytes(int elem_bytes) { return elem_bytes == 4 ? (int)sizeof(Tile<float>) : (int)sizeof(Tile<double>); }
size_t csf_pair_tiled_workspace_bytes(int64_t n_src, int64_t n_tgt, int elem_bytes) {
    if (n_src <= 0 || n_tgt <= 0) return 0;
    const int groups = elem_bytes == 4 ? tiled_plan<float>(n_src, n_tgt).n_groups : tiled_plan<double>(n_src, n_tgt).n_groups;
    return (size_t)groups * (size_t)n_tgt * 2 * (size_t)elem_bytes;
}
int csf_morton_keys_f32(const void* xycs, int64_t n, double x0, double y0, double cell, int64_t* keys, csf_stream_t st) {
    if (n <= 0) return 0;
    morton_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>((const Xycs<float>*)xycs, n, x0, y0,
                                                                                  1.0 / cell, keys);
    CSF_CHECK_LAUNCH("morton_kernel");
    return 0;
}
int csf_morton_keys_f64(const void* xycs, int64_t n, double x0, double y0, double cell, int64_t* keys, csf_stream_t st) {
    if (n <= 0) return 0;
    morton_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>((const Xycs<double>*)xycs, n, x0,
                                                                                   y0, 1.0 / cell, keys);
    CSF_CHECK_LAUNCH("morton_kernel");
    return 0;
}
int csf_tile_sources_f32(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles, csf_stream_t st) {
    return tile_sources<float>(xycs, n, perm, sorted, tiles, (cudaStream_t)st);
}
int csf_tile_sources_f64(const void* xycs, int64_t n, const int64_t* perm, void* sorted, void* tiles, csf_stream_t st) {
    return tile_sources<double>(xycs, n, perm, sorted, tiles, (cudaStream_t)st);
}
int csf_pair_forces_tiled_f32(const void* sorted, const void* tiles, int64_t n_src, const void* tgt, int64_t n_tgt,
                              const CsfFieldParams* fp, float* frep, int accumulate, void* ws, size_t wsb,
                              unsigned long long* stats, csf_stream_t st) {
    return pair_tiled<float>(sorted, tiles, n_src, tgt, n_tgt, fp, frep, accumulate, ws, wsb, stats, (cudaStream_t)st);
}
int csf_pair_forces_tiled_f64(const void* sorted, const void* tiles, int64_t n_src, const void* tgt, int64_t n_tgt,
                              const CsfFieldParams* fp, double* frep, int accumulate, void* ws, size_t wsb,
                              unsigned long long* stats, csf_stream_t st) {
    return pair_tiled<double>(sorted, tiles, n_src, tgt, n_tgt, fp, frep, accumulate, ws, wsb, stats, (cudaStream_t)st);
}

}  // extern "C"
