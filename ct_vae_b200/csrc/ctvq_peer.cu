// One-shot all-reduce of the stacked codebook gradient over NVLink peer memory (declared in include/ctvq.h).
// Replaces the NCCL call for the path's only collective when every rank sits on one NVSwitch box: the message is
// 32 KB (latency-bound), so each rank simply reads all peers' slots through P2P-mapped pointers and sums them in
// rank order after a flag handshake — one kernel, no second barrier (slots alternate by epoch parity).
#include <stdlib.h>
#include <string.h>

#include "ctvq_common.cuh"

namespace ctvq {
namespace {
struct PeerParams {
    const float* grad[CTVQ_MAX_PEERS];
    unsigned int* flags[CTVQ_MAX_PEERS];
    float* out;
    size_t count;
    int world, rank;
    unsigned int epoch;
    float scale;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(512) peer_allreduce_kernel(const PeerParams p) {
    // (1) this rank's gradient slot is complete (stream order): tell every peer
    if (blockIdx.x == 0 && threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + p.rank, p.epoch);
    // (2) wait until every peer's slot of this epoch is complete (flags are monotonic, so every CTA may poll them)
    if (threadIdx.x < p.world) {
        const unsigned int* f = p.flags[p.rank] + threadIdx.x;
        const long long t0 = clock64();
        // back off between polls: in overlap mode this kernel shares its SMs with the next forward, whose persistent CTAs
        // have a static share of the tiles -- a warp spinning at full rate on one scheduler slows the whole kernel
        while ((int)(ld_acquire_sys(f) - p.epoch) < 0) {
            __nanosleep(400);
            if (clock64() - t0 > 6000000000LL) __trap();  // ~3 s: a missing peer traps instead of hanging the GPU
        }
    }
    __syncthreads();
    // (3) sum the slots in rank order: deterministic and bit-identical on every rank
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.count; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
#pragma unroll
        for (int r = 0; r < CTVQ_MAX_PEERS; ++r)
            if (r < p.world) s += __ldcv(p.grad[r] + i);
        p.out[i] = s * p.scale;
    }
}

inline size_t flags_offset(size_t count_max) { return (2 * count_max * sizeof(float) + 255) & ~(size_t)255; }
}  // namespace
}  // namespace ctvq

using namespace ctvq;

extern "C" {

size_t ctvq_peer_buffer_bytes(size_t count_max, int world) {
    (void)world;
    return flags_offset(count_max) + 256;
}

int ctvq_peer_alloc(void** dev_ptr_out, size_t count_max, int world, int device) {
    if (!dev_ptr_out || world < 1 || world > CTVQ_MAX_PEERS || count_max == 0) return CTVQ_E_BADARG;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    const size_t bytes = ctvq_peer_buffer_bytes(count_max, world);
    e = cudaMalloc(dev_ptr_out, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*dev_ptr_out, 0, bytes);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}

int ctvq_peer_free(void* dev_ptr, int device) {
    if (!dev_ptr) return CTVQ_E_BADARG;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaFree(dev_ptr);
}

int ctvq_peer_export(void* dev_ptr, void* handle64_out, int device) {
    if (!dev_ptr || !handle64_out) return CTVQ_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == CTVQ_IPC_HANDLE_BYTES, "IPC handle size");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, dev_ptr);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle64_out, &h, sizeof(h));
    return CTVQ_OK;
}

int ctvq_peer_import(const void* handle64, void** dev_ptr_out, int device) {
    if (!handle64 || !dev_ptr_out) return CTVQ_E_BADARG;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    return (int)cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
}

int ctvq_peer_close(void* dev_ptr, int device) {
    if (!dev_ptr) return CTVQ_E_BADARG;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaIpcCloseMemHandle(dev_ptr);
}

float* ctvq_peer_slot(void* own_buf, size_t count_max, unsigned epoch) {
    return static_cast<float*>(own_buf) + (size_t)(epoch & 1u) * count_max;
}

int ctvq_peer_allreduce(void* const* peer_bufs, int world, int rank, size_t count_max, size_t count, unsigned epoch,
                        float scale, float* out, int device, void* stream) {
    if (!peer_bufs || !out || world < 1 || world > CTVQ_MAX_PEERS || rank < 0 || rank >= world || count == 0 ||
        count > count_max)
        return CTVQ_E_BADARG;
    PeerParams p;
    memset(&p, 0, sizeof(p));
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r]) return CTVQ_E_BADARG;
        p.grad[r] = static_cast<const float*>(peer_bufs[r]) + (size_t)(epoch & 1u) * count_max;
        p.flags[r] = reinterpret_cast<unsigned int*>(static_cast<char*>(peer_bufs[r]) + flags_offset(count_max));
    }
    p.out = out; p.count = count; p.world = world; p.rank = rank; p.epoch = epoch; p.scale = scale;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    // same shared-memory carve-out as the quantiser kernels (which take ~210 KB per SM): an SM configured for a small
    // carve-out would have to drain and reconfigure before it can host the next forward's CTA, i.e. in overlap mode the
    // forward would wait for this kernel's handshake on those SMs
    static bool carveout_set = false;
    if (!carveout_set) {
        cudaFuncSetAttribute(peer_allreduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carveout_set = true;
    }
    // SMALL CTAs: in overlap mode this kernel must co-reside with the next forward's 544-thread, 96-register CTA on the
    // same SMs.  Measured at N=2: 512-thread CTAs delay the forward by 9 us (0.173 vs 0.164 ms), 64-thread CTAs do not.
    static const int threads = [] { const char* e = getenv("CTVQ_PEER_THREADS"); const int t = e ? atoi(e) : 64; return t >= 32 && t <= 512 ? t : 64; }();
    size_t blocks = (count + (size_t)threads * 32 - 1) / ((size_t)threads * 32);
    if (blocks < 1) blocks = 1;
    if (blocks > 148) blocks = 148;
    static const int max_blocks = [] { const char* e = getenv("CTVQ_PEER_BLOCKS"); return e ? atoi(e) : 148; }();
    if (max_blocks >= 1 && blocks > (size_t)max_blocks) blocks = (size_t)max_blocks;
    peer_allreduce_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
    const int rc = (int)cudaGetLastError();
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
