// C-ABI entry points of libctvq.so (declared in include/ctvq.h): argument validation, kernel dispatch,
// and the dlopen()ed NCCL wrappers.  No torch types; every pointer is a caller-owned device pointer.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "ctvq_common.cuh"

namespace ctvq {
static std::atomic<int> g_path{CTVQ_PATH_AUTO};
static thread_local int t_last_path = 0;

struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

}  // namespace ctvq
namespace ctvq {
int sm_count() {
    static std::atomic<int> cache[64];  // per device ordinal; 0 = not queried yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

static int check_shape(int64_t B, int Dtot, int HW, int C, int d, int K, int cs, int dtype) {
    if (B <= 0 || Dtot <= 0 || HW <= 0 || C <= 0 || d <= 0 || K <= 0 || cs < 0) return CTVQ_E_BADARG;
    if (C > CTVQ_MAX_CODEBOOKS) return CTVQ_E_UNSUPPORTED;
    if ((int64_t)(C - 1) * cs + d > Dtot) return CTVQ_E_BADARG;
    if (dtype != CTVQ_F32 && dtype != CTVQ_BF16) return CTVQ_E_UNSUPPORTED;
    if (B * (int64_t)HW > (int64_t)1 << 40) return CTVQ_E_UNSUPPORTED;
    return CTVQ_OK;
}

static int fill(QuantParams& p, const void* const* codebooks, int64_t B, int Dtot, int HW, int C, int d, int K, int cs,
                int dtype, void* workspace, size_t ws_bytes) {
    if (!codebooks || !workspace) return CTVQ_E_BADARG;
    if (ws_bytes < sizeof(Workspace)) return CTVQ_E_WORKSPACE;
    memset(&p, 0, sizeof(p));
    for (int c = 0; c < C; ++c) {
        if (!codebooks[c]) return CTVQ_E_BADARG;
        p.E[c] = static_cast<const float*>(codebooks[c]);
    }
    Workspace* ws = static_cast<Workspace*>(workspace);
    static_assert(sizeof(Workspace) <= kScratchOffset, "workspace header");
    if (ws_bytes > kScratchOffset && !(reinterpret_cast<uintptr_t>(workspace) & 255)) {
        p.scratch = static_cast<unsigned char*>(workspace) + kScratchOffset;
        p.scratch_bytes = ws_bytes - kScratchOffset;
    }
    p.loss_acc = ws->loss_acc;
    p.ticket = &ws->ticket;
    p.err = &ws->err;
    p.B = B;
    p.N = B * (int64_t)HW;
    p.n_seg = 1;
    p.Dtot = Dtot; p.HW = HW; p.C = C; p.d = d; p.K = K; p.cs = cs;
    p.dtype = dtype;
    return CTVQ_OK;
}

static int dispatch_forward(const QuantParams& p, cudaStream_t s) {
    const int want = g_path.load();
    if (p.dtype == CTVQ_BF16) {  // bf16 latents: the specialised kind::f16 kernel where it exists, else the SIMT kernel
        if (want != CTVQ_PATH_SIMT) {
            const int rc = launch_forward_tc_bf16(p, s);
            if (rc != CTVQ_E_UNSUPPORTED) { t_last_path = CTVQ_PATH_TC; return rc; }
            if (want == CTVQ_PATH_TC || want == CTVQ_PATH_TC_STREAM) return rc;
        }
        t_last_path = CTVQ_PATH_SIMT;
        return launch_forward_simt(p, s);
    }
    if (want == CTVQ_PATH_TC_STREAM) {
        t_last_path = CTVQ_PATH_TC;
        return launch_forward_tc_stream(p, s);
    }
    if (want == CTVQ_PATH_TC || (want == CTVQ_PATH_AUTO && tc_supported(p))) {
        const int rc = launch_forward_tc(p, s);
        if (rc != CTVQ_E_UNSUPPORTED || want == CTVQ_PATH_TC) { t_last_path = CTVQ_PATH_TC; return rc; }
    }
    t_last_path = CTVQ_PATH_SIMT;
    return launch_forward_simt(p, s);
}
}  // namespace ctvq

using namespace ctvq;

extern "C" {

int ctvq_version(void) { return CTVQ_VERSION; }

const char* ctvq_strerror(int rc) {
    switch (rc) {
        case CTVQ_OK: return "ok";
        case CTVQ_E_BADARG: return "ctvq: bad argument (null pointer, non-positive size or slices exceed the channel count)";
        case CTVQ_E_UNSUPPORTED: return "ctvq: unsupported shape or dtype for this build";
        case CTVQ_E_WORKSPACE: return "ctvq: workspace too small (see ctvq_workspace_bytes)";
        case CTVQ_E_NCCL: return "ctvq: NCCL not loaded or an NCCL call failed";
        case CTVQ_E_NOT_BUILT: return "ctvq: feature not built";
        default: break;
    }
    if (rc > 0) return cudaGetErrorString(static_cast<cudaError_t>(rc));
    return "ctvq: unknown error";
}

size_t ctvq_workspace_bytes(int C, int K, int d) {
    (void)d;
    // header (zero-initialised once, self-cleaning) + scratch of the streaming single-codebook kernel (no init needed);
    // a caller that passes only the header still works: those shapes take the non-streaming kernels
    return C == 1 && K > 0 ? kScratchOffset + stream_scratch_bytes(K) : kScratchOffset;
}

int ctvq_read_and_clear_err(void* workspace, size_t ws_bytes, unsigned* err_out_host, int device, void* stream) {
    if (!workspace || !err_out_host) return CTVQ_E_BADARG;
    if (ws_bytes < sizeof(Workspace)) return CTVQ_E_WORKSPACE;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned int* err = &static_cast<Workspace*>(workspace)->err;
    cudaError_t e = cudaMemcpyAsync(err_out_host, err, sizeof(unsigned), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(err, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize(s);
}

int ctvq_set_path(int path) { return g_path.exchange(path); }
int ctvq_last_path(void) { return t_last_path; }

int ctvq_argmin(const void* const* z_segs, int n_seg, const void* const* codebooks, int64_t B, int Dtot, int HW, int C,
                int d, int K, int chan_stride, int dtype, int64_t* const* idx_out_segs,
                unsigned long long* neartie_count_out, void* workspace, size_t ws_bytes, int device, void* stream) {
    if (!z_segs || !idx_out_segs || n_seg < 1 || n_seg > CTVQ_MAX_SEGMENTS) return CTVQ_E_BADARG;
    int rc = check_shape(B, Dtot, HW, C, d, K, chan_stride, dtype);
    if (rc) return rc;
    QuantParams p;
    rc = fill(p, codebooks, B, Dtot, HW, C, d, K, chan_stride, dtype, workspace, ws_bytes);
    if (rc) return rc;
    p.n_seg = n_seg;
    for (int s = 0; s < n_seg; ++s) {
        if (!z_segs[s] || !idx_out_segs[s]) return CTVQ_E_BADARG;
        p.z[s] = static_cast<const float*>(z_segs[s]);
        p.idx[s] = reinterpret_cast<long long*>(idx_out_segs[s]);
    }
    p.fused = 0;
    p.neartie = neartie_count_out;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return dispatch_forward(p, static_cast<cudaStream_t>(stream));
}

int ctvq_gather_st_loss(const void* z, const void* const* codebooks, const int64_t* idx, int64_t B, int Dtot, int HW,
                        int C, int d, int K, int chan_stride, int dtype, float beta, void* q_out, float* loss_out,
                        void* workspace, size_t ws_bytes, int device, void* stream) {
    if (!z || !idx || !q_out || !loss_out) return CTVQ_E_BADARG;
    int rc = check_shape(B, Dtot, HW, C, d, K, chan_stride, dtype);
    if (rc) return rc;
    QuantParams p;
    rc = fill(p, codebooks, B, Dtot, HW, C, d, K, chan_stride, dtype, workspace, ws_bytes);
    if (rc) return rc;
    p.z[0] = static_cast<const float*>(z);
    p.idx[0] = reinterpret_cast<long long*>(const_cast<int64_t*>(idx));
    p.q = static_cast<float*>(q_out);
    p.loss_out = loss_out;
    p.beta = beta;
    p.fused = 1;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return launch_gather(p, static_cast<cudaStream_t>(stream));
}

int ctvq_forward(const void* z, const void* const* codebooks, int64_t B, int Dtot, int HW, int C, int d, int K,
                 int chan_stride, int dtype, float beta, int64_t* idx_out, void* q_out, float* loss_out,
                 unsigned long long* neartie_count_out, void* workspace, size_t ws_bytes, int device, void* stream) {
    if (!z || !idx_out || !q_out || !loss_out) return CTVQ_E_BADARG;
    int rc = check_shape(B, Dtot, HW, C, d, K, chan_stride, dtype);
    if (rc) return rc;
    QuantParams p;
    rc = fill(p, codebooks, B, Dtot, HW, C, d, K, chan_stride, dtype, workspace, ws_bytes);
    if (rc) return rc;
    p.z[0] = static_cast<const float*>(z);
    p.idx[0] = reinterpret_cast<long long*>(idx_out);
    p.q = static_cast<float*>(q_out);
    p.loss_out = loss_out;
    p.beta = beta;
    p.fused = 1;
    p.neartie = neartie_count_out;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return dispatch_forward(p, static_cast<cudaStream_t>(stream));
}

static int backward_impl(const void* z, const void* const* codebooks, const int64_t* idx, const void* g_out,
                         const float* g_loss, int64_t B, int Dtot, int HW, int C, int d, int K, int chan_stride, int dtype,
                         float beta, void* gz_out, float* gE_out, void* workspace, size_t ws_bytes, int device,
                         void* stream, void* const* peer_bufs, int world, int rank, size_t count_max, unsigned epoch,
                         float scale, float* gE_reduced_out) {
    if (!z || !idx || !g_loss || !gz_out || !gE_out || !codebooks || !workspace) return CTVQ_E_BADARG;
    if (ws_bytes < sizeof(Workspace)) return CTVQ_E_WORKSPACE;
    int rc = check_shape(B, Dtot, HW, C, d, K, chan_stride, dtype);
    if (rc) return rc;
    BwdParams p;
    memset(&p, 0, sizeof(p));
    for (int c = 0; c < C; ++c) {
        if (!codebooks[c]) return CTVQ_E_BADARG;
        p.E[c] = static_cast<const float*>(codebooks[c]);
    }
    p.z = static_cast<const float*>(z);
    p.idx = reinterpret_cast<const long long*>(idx);
    p.g_out = static_cast<const float*>(g_out);
    p.g_loss = g_loss;
    p.gz = static_cast<float*>(gz_out);
    p.gE = gE_out;
    p.err = &static_cast<Workspace*>(workspace)->err;
    p.B = B; p.N = B * (int64_t)HW;
    p.Dtot = Dtot; p.HW = HW; p.C = C; p.d = d; p.K = K; p.cs = chan_stride;
    p.beta = beta;
    p.dtype = dtype;
    if (peer_bufs) {  // arm the fused collective: the last CTA of whichever backward kernel runs all-reduces gE_out
        rc = make_peer_tail(p.peer, peer_bufs, world, rank, count_max, (size_t)C * K * d, epoch, scale, gE_reduced_out,
                            static_cast<Workspace*>(workspace));
        if (rc) return rc;
    }
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return launch_backward(p, static_cast<cudaStream_t>(stream));
}

int ctvq_backward(const void* z, const void* const* codebooks, const int64_t* idx, const void* g_out,
                  const float* g_loss, int64_t B, int Dtot, int HW, int C, int d, int K, int chan_stride, int dtype,
                  float beta, void* gz_out, float* gE_out, void* workspace, size_t ws_bytes, int device,
                  void* stream) {
    return backward_impl(z, codebooks, idx, g_out, g_loss, B, Dtot, HW, C, d, K, chan_stride, dtype, beta, gz_out, gE_out,
                         workspace, ws_bytes, device, stream, nullptr, 0, 0, 0, 0u, 1.0f, nullptr);
}

int ctvq_backward_allreduce(const void* z, const void* const* codebooks, const int64_t* idx, const void* g_out,
                            const float* g_loss, int64_t B, int Dtot, int HW, int C, int d, int K, int chan_stride,
                            int dtype, float beta, void* gz_out, float* gE_local, void* const* peer_bufs, int world,
                            int rank, size_t count_max, unsigned epoch, float scale, float* gE_reduced_out,
                            void* workspace, size_t ws_bytes, int device, void* stream) {
    if (!peer_bufs || !gE_reduced_out) return CTVQ_E_BADARG;
    return backward_impl(z, codebooks, idx, g_out, g_loss, B, Dtot, HW, C, d, K, chan_stride, dtype, beta, gz_out, gE_local,
                         workspace, ws_bytes, device, stream, peer_bufs, world, rank, count_max, epoch, scale,
                         gE_reduced_out);
}

int ctvq_reparam_kld_fwd(const float* mu, const float* logvar, const float* eps, int64_t B, int L, float* z_out,
                         float* kld_out, void* workspace, size_t ws_bytes, int device, void* stream) {
    if (!mu || !logvar || !eps || !z_out || !kld_out || !workspace || B <= 0 || L <= 0) return CTVQ_E_BADARG;
    if (ws_bytes < sizeof(Workspace)) return CTVQ_E_WORKSPACE;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return launch_reparam_fwd(mu, logvar, eps, B, L, z_out, kld_out, static_cast<Workspace*>(workspace),
                              static_cast<cudaStream_t>(stream));
}

int ctvq_reparam_kld_bwd(const float* mu, const float* logvar, const float* eps, const float* g_z, const float* g_kld,
                         int64_t B, int L, float* g_mu_out, float* g_logvar_out, int device, void* stream) {
    if (!mu || !logvar || !eps || !g_mu_out || !g_logvar_out || B <= 0 || L <= 0) return CTVQ_E_BADARG;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return launch_reparam_bwd(mu, logvar, eps, g_z, g_kld, B, L, g_mu_out, g_logvar_out,
                              static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------
// CT-mode codec: index <-> one-hot, one-hot cross-entropy (models/ct_mcq_vae.py:472-496, 306-311)
// ---------------------------------------------------------------------------------------------------
int ctvq_onehot_from_inds(const int64_t* idx, int64_t B, int64_t S, int K, float* onehot_out, void* workspace,
                          size_t ws_bytes, int device, void* stream) {
    if (!idx || !onehot_out || !workspace || B <= 0 || S <= 0 || K <= 0) return CTVQ_E_BADARG;
    if (ws_bytes < sizeof(Workspace)) return CTVQ_E_WORKSPACE;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return launch_onehot(reinterpret_cast<const long long*>(idx), B, S, K, onehot_out, &static_cast<Workspace*>(workspace)->err,
                         static_cast<cudaStream_t>(stream));
}

int ctvq_inds_from_onehot(const float* scores, int64_t B, int64_t S, int K, int64_t* idx_out, int device, void* stream) {
    if (!scores || !idx_out || B <= 0 || S <= 0 || K <= 0) return CTVQ_E_BADARG;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return launch_class_argmax(scores, B, S, K, reinterpret_cast<long long*>(idx_out), static_cast<cudaStream_t>(stream));
}

int ctvq_latent_ce_fwd(const float* latent, const float* latent_y, int64_t B, int64_t S, int K, int64_t* target_out,
                       float* rowsum_out, float* loss_out, void* workspace, size_t ws_bytes, int device, void* stream) {
    if (!latent || !latent_y || !target_out || !rowsum_out || !loss_out || !workspace || B <= 0 || S <= 0 || K <= 0) return CTVQ_E_BADARG;
    if (ws_bytes < sizeof(Workspace)) return CTVQ_E_WORKSPACE;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return launch_latent_ce_fwd(latent, latent_y, B, S, K, reinterpret_cast<long long*>(target_out), rowsum_out, loss_out,
                                static_cast<Workspace*>(workspace), static_cast<cudaStream_t>(stream));
}

int ctvq_latent_ce_bwd(const float* latent, const int64_t* target, const float* rowsum, const float* g_loss, int64_t B,
                       int64_t S, int K, float* g_latent_out, int device, void* stream) {
    if (!latent || !target || !rowsum || !g_loss || !g_latent_out || B <= 0 || S <= 0 || K <= 0) return CTVQ_E_BADARG;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    return launch_latent_ce_bwd(latent, reinterpret_cast<const long long*>(target), rowsum, g_loss, B, S, K, g_latent_out,
                                static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------
// NCCL (dlopen: the library has no link-time dependency on libnccl, so it loads on hosts without it)
// ---------------------------------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);
typedef int (*fn_redop_create_premulsum)(int*, void*, int, int, void*);
typedef int (*fn_redop_destroy)(int, void*);
struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_get_error_string error_string = nullptr;
    fn_redop_create_premulsum premulsum = nullptr;
    fn_redop_destroy redop_destroy = nullptr;
} g_nccl;
constexpr int kNcclFloat32 = 7;  // ncclFloat32
constexpr int kNcclSum = 0;      // ncclSum

__global__ void scale_kernel(float* x, size_t n, float s) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= s;
}
}  // namespace

int ctvq_nccl_load(const char* path) {
    if (g_nccl.handle) return CTVQ_OK;
    void* h = dlopen(path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return CTVQ_E_NCCL;
    g_nccl.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
    g_nccl.error_string = (fn_get_error_string)dlsym(h, "ncclGetErrorString");
    g_nccl.premulsum = (fn_redop_create_premulsum)dlsym(h, "ncclRedOpCreatePreMulSum");
    g_nccl.redop_destroy = (fn_redop_destroy)dlsym(h, "ncclRedOpDestroy");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_reduce) return CTVQ_E_NCCL;
    g_nccl.handle = h;
    return CTVQ_OK;
}

int ctvq_nccl_unique_id(void* id128_out) {
    if (!g_nccl.handle || !id128_out) return CTVQ_E_NCCL;
    NcclUniqueId id;
    if (g_nccl.get_unique_id(&id) != 0) return CTVQ_E_NCCL;
    memcpy(id128_out, id.internal, 128);
    return CTVQ_OK;
}

int ctvq_nccl_comm_init(void** comm_out, int nranks, int rank, const void* id128, int device) {
    if (!g_nccl.handle || !comm_out || !id128) return CTVQ_E_NCCL;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    NcclUniqueId id;
    memcpy(id.internal, id128, 128);
    const int rc = g_nccl.comm_init_rank(comm_out, nranks, id, rank);
    if (rc != 0) {
        fprintf(stderr, "ctvq: ncclCommInitRank failed: %s\n", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
        return CTVQ_E_NCCL;
    }
    return CTVQ_OK;
}

int ctvq_nccl_comm_destroy(void* comm) {
    if (!g_nccl.handle || !comm) return CTVQ_E_NCCL;
    return g_nccl.comm_destroy(comm) == 0 ? CTVQ_OK : CTVQ_E_NCCL;
}

int ctvq_allreduce_codebook_grad(void* comm, float* gE, size_t count, float scale, int device, void* stream) {
    if (!g_nccl.handle || !comm || !gE) return CTVQ_E_NCCL;
    if (count == 0) return CTVQ_OK;
    DeviceGuard g(device);
    if (g.err != cudaSuccess) return (int)g.err;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // scale folded into the reduction (sum of scale*x): ONE NCCL kernel, no separate scaling pass
    if (scale != 1.0f && g_nccl.premulsum && g_nccl.redop_destroy) {
        int op = 0;
        float sc = scale;
        if (g_nccl.premulsum(&op, &sc, kNcclFloat32, /*ncclScalarHostImmediate*/ 1, comm) == 0) {
            const int rc = g_nccl.all_reduce(gE, gE, count, kNcclFloat32, op, comm, s);
            g_nccl.redop_destroy(op, comm);
            if (rc != 0) {
                fprintf(stderr, "ctvq: ncclAllReduce failed: %s\n", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
                return CTVQ_E_NCCL;
            }
            return CTVQ_OK;
        }
    }
    const int rc = g_nccl.all_reduce(gE, gE, count, kNcclFloat32, kNcclSum, comm, s);
    if (rc != 0) {
        fprintf(stderr, "ctvq: ncclAllReduce failed: %s\n", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
        return CTVQ_E_NCCL;
    }
    if (scale != 1.0f) {
        size_t blocks = (count + 255) / 256;
        if (blocks > (size_t)sm_count() * 4) blocks = (size_t)sm_count() * 4;
        scale_kernel<<<(unsigned)blocks, 256, 0, s>>>(gE, count, scale);
        return (int)cudaGetLastError();
    }
    return CTVQ_OK;
}

}  // extern "C"
