// Single-codebook backward kernel (VectorQuantizer of configs/vq_vae.yaml, the config-4 sweep shapes): straight-through
// + commitment gradient and the codebook-gradient scatter-add for C = 1, any K whose [K, d] accumulator fits in shared
// memory.  Replaces autograd through models/vq_vae.py:43-53 (one_hot^T @ g, ~10 elementwise kernels, two permutes).
//
// The generic tiled kernel (ctvq_bwd.cu) parallelises the accumulation over (codebook x 32-channel chunk): with one
// codebook of 64 channels only 2 of its 8 warps work.  Here EVERY warp walks rows (lanes along the channels, coalesced
// 128-byte reads of the winning code from L2) and adds q - z into ONE shared [K, d] accumulator with shared-memory
// atomics (red.shared.add.f32: fire-and-forget, no read-modify-write latency chain); the same pass leaves q - z in place
// of z in the staged tile, so the grad_z pass (lanes along HW, 128-bit loads/stores) never gathers from the codebook.
// A CTA is two independent HALVES of NT/2 threads (named barriers), each on its own tile, sharing the accumulator:
// one half's global loads overlap the other's arithmetic even when the accumulator leaves room for one CTA per SM only.
// One red.global.add flush per CTA.
#include <stdlib.h>

#include "ctvq_common.cuh"

namespace ctvq {
namespace {

constexpr int kHT = 256;  // threads per half

__device__ __forceinline__ void half_sync(int half) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(kHT) : "memory");
}

// TM: rows per tile (per half); GACC: the [K, d] accumulator does not fit in shared memory -> the (channel-coalesced,
// 128 bytes per warp) atomics go straight to grad_E in L2: with thousands of codes the rows rarely collide
template <int TM, bool GACC>
__global__ void __launch_bounds__(2 * kHT) vq_bwd_c1_kernel(const BwdParams p, const int ntiles) {
    extern __shared__ __align__(16) float smem[];
    const int d = p.d, K = p.K, HW = p.HW;
    constexpr int ZS = TM + 1;  // odd stride: conflict-free with lanes along rows (staging, grad_z) and along channels (accumulate)
    const int kd = K * d;
    const int tid = threadIdx.x, half = tid / kHT, ht = tid % kHT, lane = ht & 31, hw_ = ht >> 5;  // hw_: warp within the half
    float* acc = smem;                                         // [K][d], shared by both halves (absent when GACC)
    float* zd = acc + (GACC ? 0 : kd) + (size_t)half * ((size_t)d * ZS + TM);  // [d][ZS]: z, then q - z in place
    int* idx_s = reinterpret_cast<int*>(zd + (size_t)d * ZS);  // [TM]
    if (!GACC)
        for (int i = tid; i < kd; i += 2 * kHT) acc[i] = 0.0f;
    const float gl = __ldg(p.g_loss);
    const double nd = (double)p.N * (double)d;
    const float coef_e = (float)(2.0 / nd) * gl;                    // weight of (q - z) in d vq_loss / d E
    const float coef_z = (float)(2.0 * (double)p.beta / nd) * gl;   // weight of (z - q) in d vq_loss / d z
    const float* __restrict__ E = p.E[0];
    const int jchunks = (d + 31) >> 5;
    __syncthreads();

    for (int tile = blockIdx.x * 2 + half; tile < ntiles; tile += gridDim.x * 2) {
        const long long row0 = (long long)tile * TM;
        const int mcount = (int)min((long long)TM, p.N - row0);
        half_sync(half);  // previous tile of this half fully consumed
        // ---- stage indices and z: 128-bit loads along HW, 8 in flight per thread --------------------------------------
        if (ht < TM) {
            const int m = ht;
            int k = 0;
            if (m < mcount) {
                const long long n = row0 + m;
                const long long b = n / HW;
                const long long kk = __ldg(p.idx + (size_t)b * HW + (int)(n - b * HW));
                if (kk < 0 || kk >= K) { atomicOr(p.err, 1u); k = kk < 0 ? 0 : K - 1; } else k = (int)kk;
            }
            idx_s[m] = k;
        }
        {
            constexpr int MG = TM / 4;
            const int total = d * MG;
            for (int it0 = ht; it0 < total; it0 += 8 * kHT) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int it = it0 + u * kHT;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (it < total) {
                        const int ch = it / MG, m = (it % MG) * 4;
                        if (m < mcount) {
                            const long long n = row0 + m;
                            const long long b = n / HW;
                            v[u] = __ldg(reinterpret_cast<const float4*>(p.z + ((size_t)b * d + ch) * HW + (int)(n - b * HW)));
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int it = it0 + u * kHT;
                    if (it < total) {
                        float* dst = zd + (it / MG) * ZS + (it % MG) * 4;
                        dst[0] = v[u].x; dst[1] = v[u].y; dst[2] = v[u].z; dst[3] = v[u].w;
                    }
                }
            }
        }
        half_sync(half);
        // ---- accumulate: warps over rows (4 rows' codebook reads in flight), lanes along the channels -----------------
        for (int m0 = hw_; m0 < mcount; m0 += 4 * (kHT / 32)) {
            int k[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = m0 + u * (kHT / 32);
                k[u] = m < mcount ? idx_s[m] : -1;
            }
            for (int jc = 0; jc < jchunks; ++jc) {
                const int j = jc * 32 + lane;
                if (j < d) {
                    float e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = k[u] >= 0 ? __ldg(E + (size_t)k[u] * d + j) : 0.0f;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (k[u] >= 0) {
                            const int m = m0 + u * (kHT / 32);
                            const float diff = __fsub_rn(e[u], zd[j * ZS + m]);  // q - z
                            if (GACC) atomicAdd(p.gE + (size_t)k[u] * d + j, coef_e * diff);
                            else atomicAdd(acc + (size_t)k[u] * d + j, diff);
                            zd[j * ZS + m] = diff;
                        }
                    }
                }
            }
        }
        half_sync(half);
        // ---- grad_z: lanes along HW, 128-bit --------------------------------------------------------------------------
        constexpr int MG = TM / 4;
#pragma unroll 4
        for (int it = ht; it < d * MG; it += kHT) {
            const int ch = it / MG, m = (it % MG) * 4;
            if (m >= mcount) continue;
            const long long n = row0 + m;
            const long long b = n / HW;
            const int pp = (int)(n - b * HW);
            const size_t off = ((size_t)b * d + ch) * HW + pp;
            float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.g_out) go = __ldg(reinterpret_cast<const float4*>(p.g_out + off));
            const float* dz = zd + ch * ZS + m;
            float4 g;
            g.x = go.x - coef_z * dz[0];
            g.y = go.y - coef_z * dz[1];
            g.z = go.z - coef_z * dz[2];
            g.w = go.w - coef_z * dz[3];
            *reinterpret_cast<float4*>(p.gz + off) = g;
        }
    }
    if (!GACC) {
        __syncthreads();
        for (int i = tid; i < kd; i += 2 * kHT) {
            const float v = acc[i];
            if (v != 0.0f) atomicAdd(&p.gE[i], coef_e * v);
        }
    }
    peer_tail(p.peer, p.gE);  // fused collective (no-op unless ctvq_backward_allreduce armed it)
}

template <int TM, bool GACC>
int launch_c1(const BwdParams& p, size_t sm, int per_sm, int ntiles, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(vq_bwd_c1_kernel<TM, GACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    int grid = sm_count() * per_sm;
    if (grid * 2 > ntiles) grid = (ntiles + 1) / 2;
    vq_bwd_c1_kernel<TM, GACC><<<grid, 2 * kHT, sm, s>>>(p, ntiles);
    return (int)cudaGetLastError();
}
}  // namespace

// CTVQ_E_UNSUPPORTED: not a single full-width codebook, unaligned, the accumulator does not fit, or a problem too small
// to amortise zeroing + flushing one accumulator per CTA (the caller falls through to the other kernels)
int launch_backward_c1(const BwdParams& p, cudaStream_t s) {
    if (p.C != 1 || p.d != p.Dtot || p.HW % 4 != 0) return CTVQ_E_UNSUPPORTED;
    if (p.K < 128) return CTVQ_E_UNSUPPORTED;  // few codes: the shared atomics collide (measured 1.7x slower than the ownership kernels at K=64)
    if (reinterpret_cast<uintptr_t>(p.z) & 15) return CTVQ_E_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(p.gz) & 15) || (p.g_out && (reinterpret_cast<uintptr_t>(p.g_out) & 15))) return CTVQ_E_UNSUPPORTED;
    const size_t kd = (size_t)p.K * p.d;
    auto bytes = [&](int TM, bool gacc) { return ((gacc ? 0 : kd) + 2 * ((size_t)p.d * (TM + 1) + TM)) * sizeof(float); };
    const long long nrows = p.N;
    bool gacc = false;
    int TM = 128;
    if (bytes(128, false) > 220 * 1024) TM = 64;
    // fp32 atomics on shared memory are a compare-and-swap loop on sm_100a (ATOMS.CAST.SPIN) while red.global.add.f32 is
    // native in L2: measured at 1 M rows, K=512 x D=64: 0.434 ms (shared) vs 0.331 ms (global); K=256 x D=64: 0.416 vs
    // 0.321 ms; K=256 x D=32: 0.143 vs 0.175 ms -- so wide rows with at least 256 codes go straight to grad_E as well
    const bool prefer_global = p.d >= 64 && p.K >= 256 && p.N >= 32768;
    if (bytes(TM, false) > 220 * 1024 || prefer_global) {  // accumulator too big for shared memory (or slower there)
        gacc = true;
        TM = bytes(128, true) <= 220 * 1024 ? 128 : 64;
        if (bytes(TM, true) > 220 * 1024) return CTVQ_E_UNSUPPORTED;
        if (p.K < 256) return CTVQ_E_UNSUPPORTED;  // few codes: global atomics on the same rows would serialise
    }
    const size_t sm = bytes(TM, gacc);
    int per_sm = (int)((227 * 1024) / (sm + 1024));
    if (per_sm > 2) per_sm = 2;  // 2 x 512 threads fill the SM
    if (per_sm < 1) per_sm = 1;
    const long long ntiles_ll = (nrows + TM - 1) / TM;
    if (ntiles_ll > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
    if (!gacc) {
        // every CTA zeroes and flushes a [K,d] accumulator: only worth it when the rows outweigh that
        const long long ctas = (ntiles_ll + 1) / 2 < (long long)sm_count() * per_sm ? (ntiles_ll + 1) / 2 : (long long)sm_count() * per_sm;
        if ((double)p.N * p.d < 2.0 * (double)ctas * (double)kd) return CTVQ_E_UNSUPPORTED;
    } else if (p.N < 32768) {
        return CTVQ_E_UNSUPPORTED;  // tiny batches keep the direct kernel
    }
    if (gacc) return TM == 128 ? launch_c1<128, true>(p, sm, per_sm, (int)ntiles_ll, s) : launch_c1<64, true>(p, sm, per_sm, (int)ntiles_ll, s);
    return TM == 128 ? launch_c1<128, false>(p, sm, per_sm, (int)ntiles_ll, s) : launch_c1<64, false>(p, sm, per_sm, (int)ntiles_ll, s);
}

}  // namespace ctvq
