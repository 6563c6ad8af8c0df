// Trajectory stream: the per-step record of the crowd that the reference keeps in every
// vehicle's `traj` / `trajF` arrays (vehicle.py:320-325, :1407-1413) and hands to SUMO one road user at a
// time (intersection.py:679-688).  Here one launch per step appends the state columns of every model group
// and the total forces to a device ring; the host takes a whole chunk of steps with one copy
// (cyclistsocialforce_b200/trajstream.py).
#include "csf_common.cuh"

namespace {

// Several contiguous device-to-device copies in ONE launch: blockIdx.y = segment, grid-stride over 16-byte
// words (4-byte words if a pointer or the length is not 16-byte aligned; every array of the state is a
// multiple of 4 bytes).
__global__ void __launch_bounds__(256) copy_segments_kernel(CsfCopySegments segs) {
    const CsfCopySegment sg = segs.seg[blockIdx.y];
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uintptr_t a = reinterpret_cast<uintptr_t>(sg.src) | reinterpret_cast<uintptr_t>(sg.dst) | (uintptr_t)sg.bytes;
    if ((a & 15u) == 0) {
        const uint4* __restrict__ s = static_cast<const uint4*>(sg.src);
        uint4* __restrict__ d = static_cast<uint4*>(sg.dst);
        const size_t n = (size_t)sg.bytes >> 4;
        for (size_t i = i0; i < n; i += stride) d[i] = s[i];
    } else {
        const uint32_t* __restrict__ s = static_cast<const uint32_t*>(sg.src);
        uint32_t* __restrict__ d = static_cast<uint32_t*>(sg.dst);
        const size_t n = (size_t)sg.bytes >> 2;
        for (size_t i = i0; i < n; i += stride) d[i] = s[i];
    }
}

// Road-user churn (intersection.py:458-539, :576-634): every per-agent array of a model group is re-laid in ONE
// launch.  A segment describes an array as (outer, n, inner): dst[o][dst_off + j][:] = src[o][idx ? idx[j] : j][:]
// for j < count -- rows of agents (outer = 1), column-wise arrays such as the history rings (outer = rows), with
// an index list (select / compact) or without (append the agents of another group behind dst_off).
// blockIdx.y = segment, grid-stride over 4-byte words.
__global__ void __launch_bounds__(256) gather_segments_kernel(CsfGatherSegments segs) {
    const CsfGatherSegment sg = segs.seg[blockIdx.y];
    const int64_t wpr = sg.inner_bytes >> 2;                       // words per (outer, agent) element of the source
    const int64_t wpd = (sg.dst_inner_bytes > 0 ? sg.dst_inner_bytes : sg.inner_bytes) >> 2;   // ... of the destination
    const int64_t total = sg.outer * sg.count * wpr;
    const uint32_t* __restrict__ src = static_cast<const uint32_t*>(sg.src);
    uint32_t* __restrict__ dst = static_cast<uint32_t*>(sg.dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = i % wpr, j = (i / wpr) % sg.count, o = i / (wpr * sg.count);
        const int64_t from = sg.idx ? sg.idx[j] : j;
        dst[(o * sg.n_dst + sg.dst_off + j) * wpd + w] = src[(o * sg.n_src + from) * wpr + w];
    }
}

// SFM heading (rad, counter-clockwise from +x) -> SUMO angle (deg, clockwise from north), utils.py:89-111
// angleSFMtoSUMO; and the positions, as one [n][3] double array for a batched traci.vehicle.moveToXY.
template <typename T>
__global__ void sumo_pose_kernel(const double* __restrict__ x, const double* __restrict__ y, const T* __restrict__ psi,
                                 int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double th = 90.0 - (double)psi[i] * (360.0 / CSF_TWO_PI);
    th -= floor(th / 360.0) * 360.0;
    out[3 * i] = x[i];
    out[3 * i + 1] = y[i];
    out[3 * i + 2] = th;
}

}  // namespace

extern "C" {

int csf_copy_segments(const CsfCopySegments* segs, csf_stream_t s) {
    if (segs->n <= 0) return 0;
    if (segs->n > CSF_MAX_COPY_SEGMENTS) { csf_set_error("csf_copy_segments: too many segments", cudaErrorInvalidValue); return -(int)cudaErrorInvalidValue; }
    int64_t most = 0;
    for (int i = 0; i < segs->n; ++i) {
        if (segs->seg[i].bytes & 3) { csf_set_error("csf_copy_segments: lengths must be multiples of 4 bytes", cudaErrorInvalidValue); return -(int)cudaErrorInvalidValue; }
        most = segs->seg[i].bytes > most ? segs->seg[i].bytes : most;
    }
    if (most == 0) return 0;
    const int64_t want = (most / 16 + 255) / 256;
    const unsigned gx = (unsigned)(want < 1 ? 1 : (want > 592 ? 592 : want));      // <= 4 CTAs per SM and segment
    copy_segments_kernel<<<dim3(gx, (unsigned)segs->n), 256, 0, (cudaStream_t)s>>>(*segs);
    CSF_CHECK_LAUNCH("copy_segments_kernel");
    return 0;
}

int csf_gather_segments(const CsfGatherSegments* segs, csf_stream_t s) {
    if (segs->n <= 0) return 0;
    if (segs->n > CSF_MAX_GATHER_SEGMENTS) {
        csf_set_error("csf_gather_segments: too many segments", cudaErrorInvalidValue);
        return -(int)cudaErrorInvalidValue;
    }
    int64_t most = 0;
    for (int i = 0; i < segs->n; ++i) {
        const CsfGatherSegment& g = segs->seg[i];
        if ((g.inner_bytes & 3) || g.inner_bytes <= 0 || (g.dst_inner_bytes & 3) ||
            (g.dst_inner_bytes > 0 && g.dst_inner_bytes < g.inner_bytes)) {
            csf_set_error("csf_gather_segments: element sizes must be positive multiples of 4 bytes", cudaErrorInvalidValue);
            return -(int)cudaErrorInvalidValue;
        }
        const int64_t w = g.outer * g.count * (g.inner_bytes >> 2);
        most = w > most ? w : most;
    }
    if (most == 0) return 0;
    const int64_t want = (most + 255) / 256;
    const unsigned gx = (unsigned)(want < 1 ? 1 : (want > 592 ? 592 : want));
    gather_segments_kernel<<<dim3(gx, (unsigned)segs->n), 256, 0, (cudaStream_t)s>>>(*segs);
    CSF_CHECK_LAUNCH("gather_segments_kernel");
    return 0;
}

int csf_sumo_pose_f32(const double* x, const double* y, const void* psi, int64_t n, double* out, csf_stream_t s) {
    if (n <= 0) return 0;
    sumo_pose_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(x, y, static_cast<const float*>(psi), n, out);
    CSF_CHECK_LAUNCH("sumo_pose_kernel");
    return 0;
}
int csf_sumo_pose_f64(const double* x, const double* y, const void* psi, int64_t n, double* out, csf_stream_t s) {
    if (n <= 0) return 0;
    sumo_pose_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(x, y, static_cast<const double*>(psi), n, out);
    CSF_CHECK_LAUNCH("sumo_pose_kernel");
    return 0;
}

}  // extern "C"
