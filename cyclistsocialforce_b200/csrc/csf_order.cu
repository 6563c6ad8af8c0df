// Spatial (Hilbert-curve) order of the road users, entirely on the device and inside this library:
//   csf_spatial_bbox_*   bounding box of the payload positions -> box[4] in device memory
//   csf_spatial_order_*  Hilbert keys over that box, stable radix sort of (key, index) -> perm
// The order is what makes the tiled pair kernel's bounding circles tight (csf_pair_tiled.cu); it is
// refreshed every few dozen steps.  Any order gives the same forces.  The sort itself is CUB's
// DeviceRadixSort (part of the CUDA toolkit, header-only) on 32-bit keys: four stable 8-bit passes.
#include "csf_common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {

__device__ __forceinline__ void pos_of(const Xycs<float>& e, double& x, double& y) { x = (double)e.xq; y = (double)e.yq; }
__device__ __forceinline__ void pos_of(const Xycs<double>& e, double& x, double& y) { x = e.x; y = e.y; }

// one CTA: box = {xmin, xmax, ymin, ymax} in payload units
template <typename T>
__global__ void __launch_bounds__(1024) bbox_kernel(const Xycs<T>* __restrict__ xycs, int64_t n, double* __restrict__ box) {
    double lo_x = 1e300, hi_x = -1e300, lo_y = 1e300, hi_y = -1e300;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double x, y;
        pos_of(xycs[i], x, y);
        lo_x = fmin(lo_x, x); hi_x = fmax(hi_x, x);
        lo_y = fmin(lo_y, y); hi_y = fmax(hi_y, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo_x = fmin(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
        hi_x = fmax(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        lo_y = fmin(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
        hi_y = fmax(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    __shared__ double s[32][4];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s[w][0] = lo_x; s[w][1] = hi_x; s[w][2] = lo_y; s[w][3] = hi_y; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        lo_x = lane < nw ? s[lane][0] : 1e300;
        hi_x = lane < nw ? s[lane][1] : -1e300;
        lo_y = lane < nw ? s[lane][2] : 1e300;
        hi_y = lane < nw ? s[lane][3] : -1e300;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo_x = fmin(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
            hi_x = fmax(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
            lo_y = fmin(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
            hi_y = fmax(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
        }
        if (lane == 0) { box[0] = lo_x; box[1] = hi_x; box[2] = lo_y; box[3] = hi_y; }
    }
}

// index along a 2^16 x 2^16 Hilbert curve (same curve as csf_pair_tiled.cu's 64-bit keys and synthetic.spatial_order)
__device__ __forceinline__ uint32_t hilbert32(uint32_t x, uint32_t y) {
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
__global__ void keys_iota_kernel(const Xycs<T>* __restrict__ xycs, int64_t n, const double* __restrict__ box,
                                 uint32_t* __restrict__ keys, int64_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x0 = box[0], y0 = box[2];
    const double inv_cell = 65535.0 / fmax(fmax(box[1] - box[0], box[3] - box[2]), 1e-300);
    double x, y;
    pos_of(xycs[i], x, y);
    const uint32_t kx = (uint32_t)fmin(fmax((x - x0) * inv_cell, 0.0), 65535.0);
    const uint32_t ky = (uint32_t)fmin(fmax((y - y0) * inv_cell, 0.0), 65535.0);
    keys[i] = hilbert32(kx, ky);
    idx[i] = i;
}

size_t align256(size_t b) { return (b + 255) / 256 * 256; }
size_t sort_temp_bytes(int64_t n) {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int64_t*)nullptr,
                                    (int64_t*)nullptr, (int)n, 0, 32, (cudaStream_t)0);
    return tb;
}

template <typename T>
int spatial_order(const void* xycs, int64_t n, const double* box, int64_t* perm, void* ws, size_t wsb, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n > 0x7fffffff) {
        csf_set_error("csf_spatial_order: more than 2^31 - 1 road users", cudaErrorInvalidValue);
        return -(int)cudaErrorInvalidValue;
    }
    const size_t kb = align256((size_t)n * 4), vb = align256((size_t)n * 8), tb = sort_temp_bytes(n);
    if (ws == nullptr || wsb < 2 * kb + vb + tb) {
        csf_set_error("csf_spatial_order: workspace too small", cudaErrorInvalidValue);
        return -(int)cudaErrorInvalidValue;
    }
    unsigned char* w = (unsigned char*)ws;
    uint32_t* keys_in = (uint32_t*)w;
    uint32_t* keys_out = (uint32_t*)(w + kb);
    int64_t* idx = (int64_t*)(w + 2 * kb);
    void* temp = w + 2 * kb + vb;
    keys_iota_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const Xycs<T>*)xycs, n, box, keys_in, idx);
    CSF_CHECK_LAUNCH("keys_iota_kernel");
    size_t tbytes = tb;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tbytes, keys_in, keys_out, idx, perm, (int)n, 0, 32, st);
    if (e != cudaSuccess) {
        csf_set_error("csf_spatial_order: radix sort", e);
        return -(int)e;
    }
    return 0;
}

}  // namespace

extern "C" {

size_t csf_spatial_order_workspace_bytes(int64_t n) {
    if (n <= 0) return 0;
    return 2 * align256((size_t)n * 4) + align256((size_t)n * 8) + sort_temp_bytes(n);
}
int csf_spatial_bbox_f32(const void* xycs, int64_t n, double* box_dev, csf_stream_t st) {
    if (n <= 0) return 0;
    bbox_kernel<float><<<1, 1024, 0, (cudaStream_t)st>>>((const Xycs<float>*)xycs, n, box_dev);
    CSF_CHECK_LAUNCH("bbox_kernel");
    return 0;
}
int csf_spatial_bbox_f64(const void* xycs, int64_t n, double* box_dev, csf_stream_t st) {
    if (n <= 0) return 0;
    bbox_kernel<double><<<1, 1024, 0, (cudaStream_t)st>>>((const Xycs<double>*)xycs, n, box_dev);
    CSF_CHECK_LAUNCH("bbox_kernel");
    return 0;
}
int csf_spatial_order_f32(const void* xycs, int64_t n, const double* box_dev, int64_t* perm, void* ws, size_t wsb,
                          csf_stream_t st) {
    return spatial_order<float>(xycs, n, box_dev, perm, ws, wsb, (cudaStream_t)st);
}
int csf_spatial_order_f64(const void* xycs, int64_t n, const double* box_dev, int64_t* perm, void* ws, size_t wsb,
                          csf_stream_t st) {
    return spatial_order<double>(xycs, n, box_dev, perm, ws, wsb, (cudaStream_t)st);
}

}  // extern "C"
