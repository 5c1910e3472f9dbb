// Host side of the one-shot NVLink peer-memory all-reduce (protocol and device code: ctvq_peer.cuh): symmetric-buffer
// management over CUDA IPC, the stand-alone collective kernel, and the fused backward + all-reduce entry point.
#include <stdlib.h>
#include <string.h>

#include "ctvq_common.cuh"

namespace ctvq {
namespace {
__global__ void __launch_bounds__(512) peer_allreduce_kernel(const PeerTail t, const float* src) { peer_push_reduce(t, src); }

inline size_t flags_offset(size_t count_max, int world) { return (2 * (size_t)world * count_max * sizeof(float) + 255) & ~(size_t)255; }

unsigned long long peer_timeout_ns() {  // wall-clock bound on the wait for a peer, CTVQ_PEER_TIMEOUT_MS (default 30 s)
    static const unsigned long long ns = [] {
        const char* e = getenv("CTVQ_PEER_TIMEOUT_MS");
        const long long ms = e ? atoll(e) : 30000;
        return (unsigned long long)(ms > 0 ? ms : 30000) * 1000000ull;
    }();
    return ns;
}
}  // namespace

// Fills the tail descriptor from the caller's table of mapped peer buffers.  CTVQ_E_BADARG on inconsistent arguments.
int make_peer_tail(PeerTail& t, void* const* peer_bufs, int world, int rank, size_t count_max, size_t count, unsigned epoch,
                   float scale, float* out, Workspace* ws) {
    if (!peer_bufs || !out || !ws || world < 1 || world > CTVQ_MAX_PEERS || rank < 0 || rank >= world || count == 0 ||
        count > count_max || (count_max & 3))
        return CTVQ_E_BADARG;
    memset(&t, 0, sizeof(t));
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r]) return CTVQ_E_BADARG;
        t.recv[r] = static_cast<float*>(peer_bufs[r]);
        t.flags[r] = reinterpret_cast<unsigned int*>(static_cast<char*>(peer_bufs[r]) + flags_offset(count_max, world));
    }
    t.out = out; t.ticket = &ws->ticket2; t.err = &ws->err;
    t.count = count; t.count_max = count_max; t.timeout_ns = peer_timeout_ns();
    t.world = world; t.rank = rank; t.epoch = epoch; t.scale = scale;
    return CTVQ_OK;
}
}  // namespace ctvq

using namespace ctvq;

extern "C" {

size_t ctvq_peer_buffer_bytes(size_t count_max, int world) { return flags_offset(count_max, world < 1 ? 1 : world) + 256; }

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

int ctvq_peer_allreduce(void* const* peer_bufs, int world, int rank, size_t count_max, const float* src, size_t count,
                        unsigned epoch, float scale, float* out, void* workspace, size_t ws_bytes, int device, void* stream) {
    if (!src || !workspace) return CTVQ_E_BADARG;
    if (ws_bytes < sizeof(Workspace)) return CTVQ_E_WORKSPACE;
    PeerTail t;
    const int rc0 = make_peer_tail(t, peer_bufs, world, rank, count_max, count, epoch, scale, out, static_cast<Workspace*>(workspace));
    if (rc0) return rc0;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    peer_allreduce_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(t, src);
    const int rc = (int)cudaGetLastError();
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
