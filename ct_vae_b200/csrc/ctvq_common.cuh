// Shared declarations of the ctvq kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
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
    int dtype;  // CTVQ_F32 / CTVQ_BF16: element type behind z[] and q (codebooks are always the fp32 parameters)
};

// Element I/O of latents / outputs / gradients.  bf16 mode (include/ctvq.h) is the fp32 arithmetic contract applied to
// bf16-ROUNDED operands: latents arrive as bf16 (exact in fp32), codebook values are rounded to bf16 as they are read
// (cb), every product and sum is fp32, results are rounded to bf16 only when they are stored.
template <typename T> struct IO;
template <> struct IO<float> {
    static constexpr int kDtype = CTVQ_F32;
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ void st4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ float cb(float e) { return e; }
    static __host__ __device__ __forceinline__ bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
};
template <> struct IO<__nv_bfloat16> {
    static constexpr int kDtype = CTVQ_BF16;
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) {
        return __uint_as_float((unsigned)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
    }
    static __device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
    static __device__ __forceinline__ void st4(__nv_bfloat16* p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 t;
        t.x = *reinterpret_cast<const unsigned*>(&a);
        t.y = *reinterpret_cast<const unsigned*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
    static __device__ __forceinline__ float cb(float e) { return __bfloat162float(__float2bfloat16_rn(e)); }
    static __host__ __device__ __forceinline__ bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
};

// Absolute term of the tensor-core candidate window, in units of (|z|^2 + max|e|^2): 2^-20 covers the fp32 rounding of
// the accumulation and of the distance formula itself; the extra 1e-6 widens the window by >= CTVQ_NEAR_TIE_REL * d1
// (d1 <= 2 (|z|^2 + max|e|^2)), so the exact SECOND-best code of every near-tie row is among the survivors as well and
// the near-tie count falls out of the exact re-scoring loop.
constexpr float kWinAbs = 9.5367431640625e-7f + 1.0e-6f;

// tcgen05.mma kind::tf32 TRUNCATES its fp32 operands to 10 explicit mantissa bits (measured: tools/tf32_rounding_probe.py,
// pinned by tests/test_stream_gpu.py::test_tf32_operands_are_truncated): x_hi = x (1 - dx), dx in [0, 2^-10), so every
// product is z e (1 - dz)(1 - de) with (1 - dz)(1 - de) in (c - 2^-10, c + 2^-10], c = 1 - 2^-10.  Dividing the tensor-core
// dot product by c CENTRES the error:  |dot_tf32 / c - z.e| <= (2^-10 / c) sum |z_j e_j| <= kTf32Eps |z||e|  -- half the
// uncentred bound 2^-9 |z||e|, for free: the kernels whose accumulator holds  z.e - |e|^2/2  scale the exact |e|^2 term by
// c instead (the score is then c times the true one, which only makes the window a hair wider), the others fold 1/c into
// the -2 they multiply the dot product with.  Half the window is roughly half the rows that need exact re-scoring.
constexpr float kTruncC = 1.0f - 9.765625e-4f;    // 1 - 2^-10, exact in fp32
constexpr float kTf32Eps = 1.03e-3f;              // >= 2^-10 / c = 9.775e-4 (5 % slack, as the uncentred 2.05e-3 had over 2^-9)
constexpr float kNeg2OverC = -2.0f / kTruncC;     // -2.00195...: dot product -> distance units, centred

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
    int dtype;     // CTVQ_F32 / CTVQ_BF16: element type behind z, g_out, gz (idx int64, gE fp32, codebooks fp32)
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
int launch_backward_ring(const BwdParams& p, cudaStream_t s);   // single codebook of many codes, [K,d] accumulator resident in shared memory, cp.async / TMA rings
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
int launch_forward_tc_bf16(const QuantParams& p, cudaStream_t s);  // kind::f16 kernels for bf16 latents; CTVQ_E_UNSUPPORTED otherwise
int launch_forward_tc_c1(const QuantParams& p, cudaStream_t s);    // single-codebook row-split tcgen05 kernels
int launch_forward_tc_stream(const QuantParams& p, cudaStream_t s);  // single codebook of any size streamed through a TMA ring
int launch_forward_tc_res(const QuantParams& p, cudaStream_t s);     // single codebook resident in shared memory (K <= 512 at D=64)
bool res_supported(const QuantParams& p);
bool stream_supported(const QuantParams& p);
size_t stream_scratch_bytes(int K);  // scratch the streaming kernel needs behind the Workspace header
constexpr size_t kScratchOffset = 1024;  // Workspace header rounded up

}  // namespace ctvq
