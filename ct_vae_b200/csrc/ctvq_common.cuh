// Shared declarations of the ctvq kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctvq.h"
#include "ctvq_peer.cuh"

namespace ctvq {

// Kernel-parameter block: everything by value so a launch needs no device-side pointer table
// (CUDA-graph friendly; codebooks stay separate nn.Parameters as in models/mcq_vae.py:94-97).
struct QuantParams {
    const float* z[CTVQ_MAX_SEGMENTS];
    long long* idx[CTVQ_MAX_SEGMENTS];  // int64 [B, C, HW] per segment (written by argmin, read by gather)
    const float* E[CTVQ_MAX_CODEBOOKS];
    float* q;         // [B, C*d, HW]
    float* loss_out;  // [C+1]
    double* loss_acc; // workspace [C]
    unsigned int* ticket;  // workspace
    unsigned int* err;     // workspace: bit0 = index out of range seen
    long long B;      // images per segment
    long long N;      // rows per segment = B*HW
    int n_seg, Dtot, HW, C, d, K, cs;
    int tiles_per_seg;
    float beta;
    int fused;  // 0: argmin only, 1: argmin + gather + loss
    unsigned char* scratch;  // optional per-stream scratch behind the Workspace header (256-byte aligned), may be null
    size_t scratch_bytes;
    unsigned long long* neartie;  // optional device counter the kernels ADD near-tie rows to (include/ctvq.h), may be null
};

// Absolute term of the tensor-core candidate window, in units of (|z|^2 + max|e|^2): 2^-20 covers the fp32 rounding of
// the accumulation and of the distance formula itself; the extra 1e-6 widens the window by >= CTVQ_NEAR_TIE_REL * d1
// (d1 <= 2 (|z|^2 + max|e|^2)), so the exact SECOND-best code of every near-tie row is among the survivors as well and
// the near-tie count falls out of the exact re-scoring loop.
constexpr float kWinAbs = 9.5367431640625e-7f + 1.0e-6f;

// near-tie predicate on the exact fp32 distances d1 <= d2 of the best and second-best code (d2 = +inf: single code)
__device__ __forceinline__ bool near_tie(float d1, float d2) {
    return __fsub_rn(d2, d1) <= __fmul_rn(CTVQ_NEAR_TIE_REL, fabsf(d1));  // '<=': an exact tie at distance 0 counts too
}

struct Workspace {  // layout of the caller-provided zero-initialised buffer
    double loss_acc[CTVQ_MAX_CODEBOOKS];
    double kld_acc;
    unsigned int ticket;
    unsigned int ticket2;
    unsigned int err;
    unsigned int pad;
};

__device__ __forceinline__ bool lex_better(float d1, int i1, float d2, int i2) {
    // true when (d1,i1) must replace (d2,i2): smaller value, ties -> smaller index, first NaN wins
    // (torch.argmin semantics, models/vq_vae.py:35)
    const bool n1 = d1 != d1, n2 = d2 != d2;
    if (n1 | n2) return n1 && (!n2 || i1 < i2);
    return d1 < d2 || (d1 == d2 && i1 < i2);
}

__device__ __forceinline__ float dist_f32(float zz, float ee, float dot) {
    // (|z|^2 + |e|^2) - 2 z.e : the reference's association (models/vq_vae.py:30-32), no contraction
    return __fsub_rn(__fadd_rn(zz, ee), __fmul_rn(2.0f, dot));
}

struct BwdParams {
    const float* z;
    const long long* idx;
    const float* g_out;   // may be null
    const float* g_loss;  // device scalar
    const float* E[CTVQ_MAX_CODEBOOKS];
    float* gz;  // [B, Dtot, HW]
    float* gE;  // [C, K, d] (zeroed by the launcher before the kernel)
    unsigned int* err;
    long long B, N;
    int Dtot, HW, C, d, K, cs;
    float beta;
    int smem_acc;  // 1: privatise the [C,K,d] accumulator in shared memory
    PeerTail peer; // world > 1: the last CTA all-reduces gE over NVLink peer memory (ctvq_backward_allreduce)
};

struct Workspace;
// peer-tail descriptor from the caller's table of mapped symmetric buffers (ctvq_peer.cu); CTVQ_E_BADARG when inconsistent
int make_peer_tail(PeerTail& t, void* const* peer_bufs, int world, int rank, size_t count_max, size_t count, unsigned epoch,
                   float scale, float* out, Workspace* ws);

// SMs of the CURRENT device (queried once per device, then cached): every persistent grid is sized from it
int sm_count();

// launch helpers implemented in the .cu files; return 0, a cudaError_t (>0) or a CTVQ_E_* (<0)
int launch_forward_simt(const QuantParams& p, cudaStream_t s);
int launch_gather(const QuantParams& p, cudaStream_t s);
int launch_backward(const BwdParams& p, cudaStream_t s);
int launch_backward_fast(const BwdParams& p, cudaStream_t s);   // shape-specialised; CTVQ_E_UNSUPPORTED otherwise
int launch_backward_c1(const BwdParams& p, cudaStream_t s);     // single full-width codebook, shared-atomic accumulator
int launch_backward_tiled(const BwdParams& p, cudaStream_t s);  // CTVQ_E_UNSUPPORTED when [C,K,d] exceeds shared memory
int launch_reparam_fwd(const float* mu, const float* lv, const float* eps, long long B, int L, float* z, float* kld,
                       Workspace* ws, cudaStream_t s);
int launch_reparam_bwd(const float* mu, const float* lv, const float* eps, const float* g_z, const float* g_kld,
                       long long B, int L, float* g_mu, float* g_lv, cudaStream_t s);
int launch_onehot(const long long* idx, long long B, long long S, int K, float* out, unsigned int* err, cudaStream_t s);
int launch_class_argmax(const float* x, long long B, long long S, int K, long long* idx, cudaStream_t s);
int launch_latent_ce_fwd(const float* x, const float* y, long long B, long long S, int K, long long* tgt, float* rowsum,
                         float* loss, Workspace* ws, cudaStream_t s);
int launch_latent_ce_bwd(const float* x, const long long* tgt, const float* rowsum, const float* g_loss, long long B,
                         long long S, int K, float* gx, cudaStream_t s);
int launch_forward_tc(const QuantParams& p, cudaStream_t s);  // returns CTVQ_E_UNSUPPORTED when shape not covered
bool tc_supported(const QuantParams& p);
int launch_forward_tc_fast(const QuantParams& p, cudaStream_t s);  // shape-specialised tcgen05 kernels
int launch_forward_tc_c1(const QuantParams& p, cudaStream_t s);    // single-codebook row-split tcgen05 kernels
int launch_forward_tc_stream(const QuantParams& p, cudaStream_t s);  // single codebook of any size streamed through a TMA ring
int launch_forward_tc_res(const QuantParams& p, cudaStream_t s);     // single codebook resident in shared memory (K <= 512 at D=64)
bool res_supported(const QuantParams& p);
bool stream_supported(const QuantParams& p);
size_t stream_scratch_bytes(int K);  // scratch the streaming kernel needs behind the Workspace header
constexpr size_t kScratchOffset = 1024;  // Workspace header rounded up

}  // namespace ctvq
