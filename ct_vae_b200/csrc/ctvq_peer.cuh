// One-shot all-reduce of the stacked codebook gradient over NVLink peer memory, as a TAIL that any backward kernel runs
// in its LAST CTA (fused compute + collective: no second launch, no second stream), and as a stand-alone kernel for
// callers that already hold the local partial gradient.  Replaces the share of Lightning's DDP bucket all-reduce that
// carries vq_layer.*.embedding.weight.grad (run.py:99) when all ranks sit on one NVSwitch box.
//
// PUSH protocol (remote traffic is posted stores only -- nobody waits on a remote load):
//   every rank owns a symmetric buffer  recv[2 parities][world][count_max] floats + flags[world]  that its peers map
//   through CUDA IPC.  After the local partial gradient is complete, the last CTA
//     (1) copies it into slot [epoch parity][my rank] of EVERY rank's buffer (128-bit stores over NVLink; its own copy is
//         a local store),
//     (2) fences system-wide and posts `epoch` into flags[my rank] of every rank (st.release.sys),
//     (3) waits until its own flags show `epoch` for every rank (ld.acquire.sys on LOCAL memory),
//     (4) sums the `world` local slots in rank order, times `scale`, into `out` -- bit-identical on every rank.
//   Slots alternate by epoch parity, so no second barrier is needed: a rank can only push epoch e+2 after it has completed
//   all-reduce e+1, which every peer enters (posts) only after finishing its own all-reduce e.
// A peer that never arrives does not hang the GPU and does not kill the context: after `timeout_ns` of wall clock
// (%globaltimer) the waiting rank sets bit 1 of the workspace error word and finishes with what it has; the host surfaces
// it as a RuntimeError at its next check point (ctvq_read_and_clear_err).
#pragma once
#include <stdint.h>

#include "../../include/ctvq.h"

namespace ctvq {

struct PeerTail {
    float* recv[CTVQ_MAX_PEERS];          // peer r's buffer base (recv[rank] = this rank's own buffer)
    unsigned int* flags[CTVQ_MAX_PEERS];  // peer r's flag row
    float* out;                           // reduced gradient [count]
    unsigned int* ticket;                 // workspace word counting finished CTAs (self-cleaning)
    unsigned int* err;                    // workspace error word (bit 1 = peer timeout)
    unsigned long long count, count_max;
    unsigned long long timeout_ns;
    int world, rank;                      // world == 0: tail inactive (world == 1 is a valid one-rank collective)
    unsigned int epoch;
    float scale;
};

__device__ __forceinline__ unsigned int peer_ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Steps (1)-(4) above, executed by ONE CTA (all of its threads).  `src`: the complete local partial gradient [count]
// (device-coherent reads: it was accumulated with red.global.add by many CTAs).
__device__ __forceinline__ void peer_push_reduce(const PeerTail& t, const float* src) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t slot = ((size_t)(t.epoch & 1u) * t.world + t.rank) * t.count_max;
    const size_t n4 = t.count >> 2;
    // (1) push (register-light on purpose: this tail is inlined into every backward kernel and must not raise their
    // register count -- an unrolled version cost the single-codebook kernels their second CTA per SM; the loads are L2 hits)
    for (size_t i = tid; i < n4; i += nt) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + i);
#pragma unroll
        for (int r = 0; r < CTVQ_MAX_PEERS; ++r)
            if (r < t.world) reinterpret_cast<float4*>(t.recv[r] + slot)[i] = v;
    }
    for (size_t i = (n4 << 2) + tid; i < t.count; i += nt) {
        const float v = __ldcg(src + i);
        for (int r = 0; r < t.world; ++r) t.recv[r][slot + i] = v;
    }
    // (2) post: every thread's pushes are ordered system-wide before the flag stores
    __threadfence_system();
    __syncthreads();
    if (tid < t.world) peer_st_release_sys(t.flags[tid] + t.rank, t.epoch);
    // (3) wait for every rank's push into OUR buffer (flags are monotonic epochs; local polling)
    if (tid < t.world) {
        const unsigned int* f = t.flags[t.rank] + tid;
        const unsigned long long t0 = peer_globaltimer();
        while ((int)(peer_ld_acquire_sys(f) - t.epoch) < 0) {
            __nanosleep(200);
            if (peer_globaltimer() - t0 > t.timeout_ns) { atomicOr(t.err, 2u); break; }
        }
    }
    __syncthreads();
    // (4) rank-ordered sum of the local slots
    const float* base = t.recv[t.rank] + (size_t)(t.epoch & 1u) * t.world * t.count_max;
    for (size_t i = tid; i < n4; i += nt) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < CTVQ_MAX_PEERS; ++r)  // rank order: the same association on every rank
            if (r < t.world) {
                const float4 v = __ldcv(reinterpret_cast<const float4*>(base + (size_t)r * t.count_max) + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        reinterpret_cast<float4*>(t.out)[i] = make_float4(s.x * t.scale, s.y * t.scale, s.z * t.scale, s.w * t.scale);
    }
    for (size_t i = (n4 << 2) + tid; i < t.count; i += nt) {
        float s = 0.0f;
        for (int r = 0; r < t.world; ++r) s += __ldcv(base + (size_t)r * t.count_max + i);
        t.out[i] = s * t.scale;
    }
}

// Tail of a backward kernel: call from EVERY thread of EVERY CTA after the CTA's last write to the local partial gradient
// `src` (gE_out).  The last CTA to arrive runs the collective; the others return immediately (no co-residency assumption).
__device__ __forceinline__ void peer_tail(const PeerTail& t, const float* src) {
    if (t.world < 1) return;
    __shared__ unsigned int s_peer_last;
    __threadfence();  // this CTA's red.global.adds are visible device-wide before its ticket
    __syncthreads();
    if (threadIdx.x == 0) s_peer_last = (atomicAdd(t.ticket, 1u) == gridDim.x * gridDim.y - 1u);
    __syncthreads();
    if (!s_peer_last) return;
    if (threadIdx.x == 0) *t.ticket = 0u;
    __threadfence();
    peer_push_reduce(t, src);
}

}  // namespace ctvq
