// Payload exchange between the GPUs of one node over NVLink peer memory (no NCCL on the step path).
//
// One crowd is sharded by contiguous agent range (distributed.py); after K2/K3 every rank has new
// payload entries (16 B / 32 B per agent) for its own range and every other rank needs them before its
// next K1.  Each rank owns one peer-mapped buffer (cudaMalloc + CUDA IPC): the full payload array plus
// two arrays of monotonic flags.  An exchange epoch e = (pushes completed) + 1; per step, in stream order:
//   wait data    : spin until every peer's push of epoch e - 1 has landed in MY buffer
//   ... tile build, K1 (the only readers of the payload) ...
//   signal read  : store e into every peer's read_flags[me]   ("I am done reading your entries")
//   ... K2/K3 writes my own range locally ...
//   push         : wait until every peer has signalled read e, store my range into every peer's buffer
//                  (16-byte NVLink stores), system fence, pushes completed = e, store e into every peer's
//                  data_flags[me]
// The production step embeds the three in its own kernels (csf_peer.cuh: the tile-build kernel waits,
// the per-agent kernel signals, steps the agents and pushes); the stand-alone kernels below do the same
// for a payload written by something else (Engine.pack after a host upload, the first exchange).  All
// ranks issue the same sequence, so the counters agree without any host round trip and every kernel has
// fixed arguments (they replay from a CUDA graph).  Spins are bounded (~4 s); a time-out is sticky: the
// status word stays set, later waits and pushes are skipped and the host raises at its next step.
#include "csf_peer.cuh"
#include <string.h>

namespace {

__global__ void peer_wait_data_kernel(CsfPeerComm c) { csf_peer_wait_all(c); }

// epoch of the exchange in progress = pushes completed + 1 (every rank signals once and pushes once per epoch)
__global__ void peer_signal_read_kernel(CsfPeerComm c) {
    const uint32_t e = ld_sys(c.seq + SEQ_PUSH) + 1u;
    const int p = threadIdx.x;
    if (p < c.world && p != c.rank) st_sys(c.read_flags[p] + c.rank, e);
}

// first16 / n16: this rank's range of the payload in 16-byte units
__global__ void peer_push_kernel(CsfPeerComm c, int64_t first16, int64_t n16) {
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) s_epoch = ld_sys(c.seq + SEQ_PUSH) + 1u;
    __syncthreads();
    if (threadIdx.x < c.world && threadIdx.x != c.rank && peer_ok(c)) {
        if (!spin_ge(c.read_flags[c.rank] + threadIdx.x, s_epoch)) peer_fail(c, 2u);
    }
    __syncthreads();
    if (!peer_ok(c)) return;        // out of step with the peers: nothing is pushed any more
    const uint4* src = reinterpret_cast<const uint4*>(c.payload[c.rank]) + first16;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (int p = 0; p < c.world; ++p)
            if (p != c.rank) (reinterpret_cast<uint4*>(c.payload[p]) + first16)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t s_last, s_seq;
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(c.seq + SEQ_BLOCKS, 1u);
        s_last = (t == gridDim.x - 1) ? 1u : 0u;
        if (s_last) {
            c.seq[SEQ_BLOCKS] = 0;
            s_seq = s_epoch;
            c.seq[SEQ_PUSH] = s_seq;
            __threadfence_system();
        }
    }
    __syncthreads();
    if (s_last && threadIdx.x < c.world && threadIdx.x != c.rank) st_sys(c.data_flags[threadIdx.x] + c.rank, s_seq);
}

}  // namespace

extern "C" {

int csf_peer_alloc(size_t bytes, void** devptr, void* ipc_handle_out) {
    cudaError_t e = cudaMalloc(devptr, bytes);
    if (e != cudaSuccess) { csf_set_error("csf_peer_alloc: cudaMalloc", e); return -(int)e; }
    e = cudaMemset(*devptr, 0, bytes);
    if (e != cudaSuccess) { csf_set_error("csf_peer_alloc: cudaMemset", e); return -(int)e; }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, *devptr);
    if (e != cudaSuccess) { csf_set_error("csf_peer_alloc: cudaIpcGetMemHandle", e); return -(int)e; }
    memcpy(ipc_handle_out, &h, sizeof(h));
    return 0;
}
int csf_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }
int csf_peer_open(const void* ipc_handle, void** devptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(devptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { csf_set_error("csf_peer_open: cudaIpcOpenMemHandle", e); return -(int)e; }
    return 0;
}
int csf_peer_close(void* devptr) {
    cudaError_t e = cudaIpcCloseMemHandle(devptr);
    if (e != cudaSuccess) { csf_set_error("csf_peer_close", e); return -(int)e; }
    return 0;
}
int csf_peer_free(void* devptr) {
    cudaError_t e = cudaFree(devptr);
    if (e != cudaSuccess) { csf_set_error("csf_peer_free", e); return -(int)e; }
    return 0;
}
int csf_peer_wait_data(const CsfPeerComm* c, csf_stream_t st) {
    if (c->world <= 1) return 0;
    peer_wait_data_kernel<<<1, 32, 0, (cudaStream_t)st>>>(*c);
    CSF_CHECK_LAUNCH("peer_wait_data_kernel");
    return 0;
}
int csf_peer_signal_read(const CsfPeerComm* c, csf_stream_t st) {
    if (c->world <= 1) return 0;
    peer_signal_read_kernel<<<1, 32, 0, (cudaStream_t)st>>>(*c);
    CSF_CHECK_LAUNCH("peer_signal_read_kernel");
    return 0;
}
int csf_peer_push(const CsfPeerComm* c, int64_t first_elem, int64_t n_elem, int elem_bytes, csf_stream_t st) {
    if (c->world <= 1) return 0;
    const int64_t per = elem_bytes / 16, n16 = n_elem * per;
    int64_t blocks = (n16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 4 * csf_sm_count()) blocks = 4 * csf_sm_count();
    peer_push_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)st>>>(*c, first_elem * per, n16);
    CSF_CHECK_LAUNCH("peer_push_kernel");
    return 0;
}

}  // extern "C"
