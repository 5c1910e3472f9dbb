// Index <-> one-hot converters and the one-hot cross-entropy that sit either side of the quantiser in CT mode
// (SURVEY.md §8f rank 1).  Replaces CTMCQVAE.ct_preprocess / ct_postprocess (models/ct_mcq_vae.py:472-496) and
// latent_CrossEntropy_loss (models/ct_mcq_vae.py:306-311).
//
// All tensors are seen as [B, K, S] fp32 with S = C*H*W contiguous (the reference's [B, N, K*H, W] one-hot layout after
// its permute, made contiguous) and indices as [B, S] int64 (= [B, C, H, W]).  Threads run along S (coalesced, 128-bit
// where alignment allows), the K classes are a strided loop: pure HBM streaming, 4*K bytes per row either way.
#include <math_constants.h>

#include "ctvq_common.cuh"

namespace ctvq {
namespace {

// idx[b,s] -> out[b,k,s] = (idx == k).  One thread per 4 consecutive s (VEC=4) or per s (VEC=1).
template <int VEC>
__global__ void __launch_bounds__(256) onehot_kernel(const long long* __restrict__ idx, float* __restrict__ out, long long B,
                                                     long long S, int K, unsigned int* err) {
    const long long groups = S / VEC;
    const long long total = B * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / groups, s = (i - b * groups) * VEC;
        long long k[VEC];
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            k[u] = __ldg(idx + b * S + s + u);
            if (k[u] < 0 || k[u] >= K) atomicOr(err, 1u);  // F.one_hot raises (models/ct_mcq_vae.py:480); flagged, row left all-zero
        }
        float* dst = out + (size_t)b * K * S + s;
        for (int c = 0; c < K; ++c) {
            if (VEC == 4) {
                *reinterpret_cast<float4*>(dst + (size_t)c * S) = make_float4(k[0] == c ? 1.0f : 0.0f, k[1 % VEC] == c ? 1.0f : 0.0f,
                                                                               k[2 % VEC] == c ? 1.0f : 0.0f, k[3 % VEC] == c ? 1.0f : 0.0f);
            } else {
                dst[(size_t)c * S] = k[0] == c ? 1.0f : 0.0f;
            }
        }
    }
}

// torch.argmax over the class dimension: first maximum wins, the first NaN wins over any number
__device__ __forceinline__ void argmax_step(float v, int c, float& best, int& bi) {
    if (v > best || (v != v && best == best)) { best = v; bi = c; }
}

template <int VEC>
__global__ void __launch_bounds__(256) argmax_kernel(const float* __restrict__ x, long long* __restrict__ idx, long long B,
                                                     long long S, int K) {
    const long long groups = S / VEC;
    const long long total = B * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / groups, s = (i - b * groups) * VEC;
        const float* src = x + (size_t)b * K * S + s;
        float best[VEC];
        int bi[VEC];
        if (VEC == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src));
            best[0] = v.x; best[1 % VEC] = v.y; best[2 % VEC] = v.z; best[3 % VEC] = v.w;
        } else {
            best[0] = __ldg(src);
        }
#pragma unroll
        for (int u = 0; u < VEC; ++u) bi[u] = 0;
#pragma unroll 4
        for (int c = 1; c < K; ++c) {
            if (VEC == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)c * S));
                argmax_step(v.x, c, best[0], bi[0]); argmax_step(v.y, c, best[1 % VEC], bi[1 % VEC]);
                argmax_step(v.z, c, best[2 % VEC], bi[2 % VEC]); argmax_step(v.w, c, best[3 % VEC], bi[3 % VEC]);
            } else {
                argmax_step(__ldg(src + (size_t)c * S), c, best[0], bi[0]);
            }
        }
#pragma unroll
        for (int u = 0; u < VEC; ++u) idx[b * S + s + u] = bi[u];
    }
}

// latent_CrossEntropy_loss forward: per row  log(sum_k x'_k) - log(x'_t),  x' = max(x, 1e-4),  t = argmax_k y;  mean over
// the B*S rows.  Writes the targets (for the backward) and the per-row sum of x'.  VEC rows per thread (128-bit loads).
__device__ __forceinline__ float clamp_min_keep_nan(float x) { return x != x ? x : fmaxf(x, 1e-4f); }  // torch.clamp keeps NaN

template <int VEC>
__global__ void __launch_bounds__(256) latent_ce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            long long B, long long S, int K, long long* __restrict__ tgt,
                                                            float* __restrict__ rowsum, float* loss_out, double* acc,
                                                            unsigned int* ticket) {
    __shared__ double red[32];
    double part = 0.0;
    const long long groups = S / VEC;
    const long long total = B * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / groups, s = (i - b * groups) * VEC;
        const float* xs = x + (size_t)b * K * S + s;
        const float* ys = y + (size_t)b * K * S + s;
        float best[VEC], sum[VEC], xt[VEC];
        int t[VEC];
#pragma unroll
        for (int u = 0; u < VEC; ++u) { best[u] = 0.0f; sum[u] = 0.0f; xt[u] = 0.0f; t[u] = 0; }
#pragma unroll 4
        for (int c = 0; c < K; ++c) {
            float yv[VEC], xr[VEC];
            if (VEC == 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(ys + (size_t)c * S));
                const float4 q = __ldg(reinterpret_cast<const float4*>(xs + (size_t)c * S));
                yv[0] = a.x; yv[1 % VEC] = a.y; yv[2 % VEC] = a.z; yv[3 % VEC] = a.w;
                xr[0] = q.x; xr[1 % VEC] = q.y; xr[2 % VEC] = q.z; xr[3 % VEC] = q.w;
            } else {
                yv[0] = __ldg(ys + (size_t)c * S);
                xr[0] = __ldg(xs + (size_t)c * S);
            }
#pragma unroll
            for (int u = 0; u < VEC; ++u) {
                const float xv = clamp_min_keep_nan(xr[u]);
                sum[u] += xv;
                // torch.argmax: first maximum wins, the first NaN wins over any number
                if (c == 0 || yv[u] > best[u] || (yv[u] != yv[u] && best[u] == best[u])) { best[u] = yv[u]; t[u] = c; xt[u] = xv; }
            }
        }
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            tgt[b * S + s + u] = t[u];
            rowsum[b * S + s + u] = sum[u];
            part += (double)(logf(sum[u]) - logf(xt[u]));
        }
    }
    // block reduce -> fp64 atomic -> last block finalises
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (warp == 0) {
        double v = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            atomicAdd(acc, v);
            __threadfence();
            if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
                __threadfence();
                *loss_out = (float)(__ldcg(acc) / (double)(B * S));
                *acc = 0.0;
                *ticket = 0u;
                __threadfence();
            }
        }
    }
}

// d loss / d x[b,k,s] = g/R * [x >= 1e-4] * (1/sum' - [k == t]/x'_t)      (clamp(min) passes gradient where x >= min)
template <int VEC>
__global__ void __launch_bounds__(256) latent_ce_bwd_kernel(const float* __restrict__ x, const long long* __restrict__ tgt,
                                                            const float* __restrict__ rowsum, const float* __restrict__ g_loss,
                                                            long long B, long long S, int K, float* __restrict__ gx) {
    const long long groups = S / VEC;
    const long long total = B * groups;
    const float g = __ldg(g_loss) / (float)(B * S);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / groups, s = (i - b * groups) * VEC;
        const float* xs = x + (size_t)b * K * S + s;
        float* gs = gx + (size_t)b * K * S + s;
        int t[VEC];
        float inv[VEC];
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            t[u] = (int)__ldg(tgt + b * S + s + u);
            inv[u] = 1.0f / __ldg(rowsum + b * S + s + u);
        }
#pragma unroll 4
        for (int c = 0; c < K; ++c) {
            float xr[VEC], o[VEC];
            if (VEC == 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(xs + (size_t)c * S));
                xr[0] = q.x; xr[1 % VEC] = q.y; xr[2 % VEC] = q.z; xr[3 % VEC] = q.w;
            } else {
                xr[0] = __ldg(xs + (size_t)c * S);
            }
#pragma unroll
            for (int u = 0; u < VEC; ++u) {
                float d = inv[u];
                if (c == t[u]) d -= 1.0f / fmaxf(xr[u], 1e-4f);
                o[u] = xr[u] >= 1e-4f ? g * d : 0.0f;
            }
            if (VEC == 4) *reinterpret_cast<float4*>(gs + (size_t)c * S) = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
            else gs[(size_t)c * S] = o[0];
        }
    }
}

inline unsigned grid_for(long long items) {
    long long blocks = (items + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    return (unsigned)blocks;
}
}  // namespace

int launch_onehot(const long long* idx, long long B, long long S, int K, float* out, unsigned int* err, cudaStream_t s) {
    const bool vec = (S % 4 == 0) && !(reinterpret_cast<uintptr_t>(out) & 15);
    if (vec) onehot_kernel<4><<<grid_for(B * (S / 4)), 256, 0, s>>>(idx, out, B, S, K, err);
    else onehot_kernel<1><<<grid_for(B * S), 256, 0, s>>>(idx, out, B, S, K, err);
    return (int)cudaGetLastError();
}

int launch_class_argmax(const float* x, long long B, long long S, int K, long long* idx, cudaStream_t s) {
    const bool vec = (S % 4 == 0) && !(reinterpret_cast<uintptr_t>(x) & 15);
    if (vec) argmax_kernel<4><<<grid_for(B * (S / 4)), 256, 0, s>>>(x, idx, B, S, K);
    else argmax_kernel<1><<<grid_for(B * S), 256, 0, s>>>(x, idx, B, S, K);
    return (int)cudaGetLastError();
}

int launch_latent_ce_fwd(const float* x, const float* y, long long B, long long S, int K, long long* tgt, float* rowsum,
                         float* loss, Workspace* ws, cudaStream_t s) {
    const bool vec = (S % 4 == 0) && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15);
    if (vec) latent_ce_fwd_kernel<4><<<grid_for(B * (S / 4)), 256, 0, s>>>(x, y, B, S, K, tgt, rowsum, loss, &ws->kld_acc, &ws->ticket2);
    else latent_ce_fwd_kernel<1><<<grid_for(B * S), 256, 0, s>>>(x, y, B, S, K, tgt, rowsum, loss, &ws->kld_acc, &ws->ticket2);
    return (int)cudaGetLastError();
}

int launch_latent_ce_bwd(const float* x, const long long* tgt, const float* rowsum, const float* g_loss, long long B,
                         long long S, int K, float* gx, cudaStream_t s) {
    const bool vec = (S % 4 == 0) && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gx)) & 15);
    if (vec) latent_ce_bwd_kernel<4><<<grid_for(B * (S / 4)), 256, 0, s>>>(x, tgt, rowsum, g_loss, B, S, K, gx);
    else latent_ce_bwd_kernel<1><<<grid_for(B * S), 256, 0, s>>>(x, tgt, rowsum, g_loss, B, S, K, gx);
    return (int)cudaGetLastError();
}

}  // namespace ctvq
