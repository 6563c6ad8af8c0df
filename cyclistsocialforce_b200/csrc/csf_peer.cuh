// Device side of the payload exchange over NVLink peer memory (csf_peer.cu has the protocol): flag
// words with system scope, bounded spins, and the three steps in a form that other kernels can embed
// -- the step's first kernel waits for the peers' pushes, its last kernel signals "payload read",
// waits for the peers' signals and pushes its own range.
#pragma once
#include "csf_common.cuh"

namespace {

enum { SEQ_PUSH = 0, SEQ_READ = 1, SEQ_BLOCKS = 2, SEQ_STATUS = 3 };

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until *flag >= want; false on timeout (~4 s)
__device__ __forceinline__ bool spin_ge(const uint32_t* flag, uint32_t want) {
    if ((int32_t)(ld_sys(flag) - want) >= 0) return true;
    const unsigned long long t0 = gtimer();
    while ((int32_t)(ld_sys(flag) - want) < 0) {
        if (gtimer() - t0 > 4000000000ull) return false;
        __nanosleep(100);
    }
    return true;
}
// A wait that timed out is sticky: the status word stays set, every later wait and push is skipped (the
// ranks are out of step, nothing exchanged from here on is meaningful) and the host-mapped mirror makes
// the next Engine.step() raise.
__device__ __forceinline__ void peer_fail(const CsfPeerComm& c, uint32_t bit) {
    atomicOr(c.seq + SEQ_STATUS, bit);
    if (c.status_host != nullptr) *reinterpret_cast<volatile int32_t*>(c.status_host) = 1;
}
__device__ __forceinline__ bool peer_ok(const CsfPeerComm& c) { return ld_sys(c.seq + SEQ_STATUS) == 0; }

// Every peer's push of the current epoch has landed in MY buffer.  Threads p < world of the calling block
// wait; the caller synchronises the block afterwards.
__device__ __forceinline__ void csf_peer_wait_all(const CsfPeerComm& c) {
    const int p = threadIdx.x;
    if (p < c.world && p != c.rank && peer_ok(c)) {
        const uint32_t want = ld_sys(c.seq + SEQ_PUSH);
        if (!spin_ge(c.data_flags[c.rank] + p, want)) peer_fail(c, 1u);
    }
    __threadfence_system();
}

}  // namespace
