// SIMT (CUDA-core) kernels of the quantiser path for sm_100a:
//   vq_fwd_simt_kernel    distances + argmin (+ fused gather / straight-through / loss) with the z tile and
//                         the codebook tile staged in shared memory; the N x K distance matrix lives in
//                         registers only.  Replaces models/vq_vae.py:25-55, models/mcq_vae.py:26-64,100-127.
//   gather_st_loss_kernel gather by caller-supplied indices (models/mcq_vae.py:41-64) with 128-bit accesses.
//   backward_kernel       straight-through + commitment gradient and codebook-gradient scatter-add
//                         (autograd of models/vq_vae.py:43-53; MCQ slice overlap of models/mcq_vae.py:117).
//   reparam_kld_*         models/vanilla_vae.py:115-117,143 fused.
#include <math_constants.h>

#include "ctvq_common.cuh"

namespace ctvq {

// ------------------------------------------------------------------------------------------------
// block-level helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Adds the block's sum of `v` to *acc (one fp64 atomic per block).  All threads must call.
__device__ __forceinline__ void block_accumulate(double v, double* acc, double* red /* >= 32 doubles smem */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) atomicAdd(acc, t);
    }
}

// Returns true in every thread of exactly one block: the last one to arrive at `ticket`.
__device__ __forceinline__ bool last_block(unsigned int* ticket, unsigned int total) {
    __shared__ unsigned int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == total - 1u);
    __syncthreads();
    return s_last != 0;
}

// loss_c = m*beta + m ; total = ((0 + l0) + l1) + ...   (models/vq_vae.py:50, models/mcq_vae.py:125)
__device__ __forceinline__ void finalize_losses(const QuantParams& p) {
    __threadfence();
    if (threadIdx.x == 0) {
        float total = 0.0f;
        const double denom = (double)p.N * (double)p.d;
        for (int c = 0; c < p.C; ++c) {
            const double s = __ldcg(&p.loss_acc[c]);
            const float m = (float)(s / denom);
            const float l = __fadd_rn(__fmul_rn(m, p.beta), m);
            p.loss_out[c] = l;
            total = __fadd_rn(total, l);
            p.loss_acc[c] = 0.0;
        }
        p.loss_out[p.C] = total;
        *p.ticket = 0u;
        __threadfence();
    }
}

// ------------------------------------------------------------------------------------------------
// forward: distances + argmin (+ gather/ST/loss)
// ------------------------------------------------------------------------------------------------
constexpr int kTK = 64;       // codes per tile
constexpr int kDJ = 32;       // channels per staged codebook chunk
constexpr int kES = kDJ + 4;  // padded row stride of the codebook chunk (16B-group conflict-free LDS.128)

template <int TM, typename T>
__global__ void __launch_bounds__(256) vq_fwd_simt_kernel(const QuantParams p) {
    constexpr int RM = TM / 16;  // rows per thread
    extern __shared__ __align__(16) float smem[];
    const int d = p.d, K = p.K, HW = p.HW;
    const int dchunks = (d + kDJ - 1) / kDJ;
    const int dpad = dchunks * kDJ;
    float* zs = smem;                    // [dpad][TM]  channel-major, row contiguous
    float* es = zs + (size_t)dpad * TM;  // [kTK][kES]
    float* zz_s = es + kTK * kES;        // [TM]
    float* ee_s = zz_s + TM;             // [kTK]
    int* idx_s = reinterpret_cast<int*>(ee_s + kTK);  // [TM]
    __shared__ double red[32];

    const int tid = threadIdx.x;
    const int c = blockIdx.y;
    const int seg = blockIdx.x / p.tiles_per_seg;
    const long long row0 = (long long)(blockIdx.x - seg * p.tiles_per_seg) * TM;
    const T* __restrict__ z = reinterpret_cast<const T*>(p.z[seg]);
    const float* __restrict__ E = p.E[c];

    // ---- stage the z tile: thread owns one row m, strides over channels (coalesced along HW) ----------
    const int lm = tid % TM;
    const long long ln = row0 + lm;
    const bool lvalid = ln < p.N;
    const long long lb = lvalid ? ln / HW : 0;
    const int lp = lvalid ? (int)(ln - lb * HW) : 0;
    {
        const T* src = z + ((size_t)lb * p.Dtot + (size_t)c * p.cs) * HW + lp;
        for (int j = tid / TM; j < dpad; j += 256 / TM)
            zs[(size_t)j * TM + lm] = (lvalid && j < d) ? IO<T>::ld(src + (size_t)j * HW) : 0.0f;
    }
    __syncthreads();
    if (tid < TM) {  // |z|^2, sequential FMA chain over ascending channel (arithmetic contract)
        float a = 0.0f;
        for (int j = 0; j < d; ++j) { const float v = zs[(size_t)j * TM + tid]; a = fmaf(v, v, a); }
        zz_s[tid] = a;
    }

    const int tk = tid & 15, tm = tid >> 4;
    float bestv[RM], secv[RM];  // running best / second-best distance (the latter only feeds the near-tie count)
    int besti[RM];
#pragma unroll
    for (int r = 0; r < RM; ++r) { bestv[r] = CUDART_INF_F; secv[r] = CUDART_INF_F; besti[r] = 0; }

    const bool e_vec = ((reinterpret_cast<uintptr_t>(E) & 15) == 0) && ((d & 3) == 0);
    const int ktiles = (K + kTK - 1) / kTK;
    for (int kt = 0; kt < ktiles; ++kt) {
        float acc[RM][4];
#pragma unroll
        for (int r = 0; r < RM; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = 0.0f;
        float ee_acc = 0.0f;
        for (int jc = 0; jc < dchunks; ++jc) {
            __syncthreads();  // previous chunk fully consumed
            // stage codebook chunk es[kk][jj] = E[kt*TK+kk][jc*DJ+jj], zero-filled outside [K) x [d)
            if (e_vec) {
                for (int e = tid; e < kTK * (kDJ / 4); e += 256) {
                    const int kk = e / (kDJ / 4), jq = e % (kDJ / 4);
                    const int k = kt * kTK + kk, j = jc * kDJ + jq * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k < K && j < d) v = __ldg(reinterpret_cast<const float4*>(E + (size_t)k * d + j));
                    v = make_float4(IO<T>::cb(v.x), IO<T>::cb(v.y), IO<T>::cb(v.z), IO<T>::cb(v.w));
                    *reinterpret_cast<float4*>(es + kk * kES + jq * 4) = v;
                }
            } else {
                for (int e = tid; e < kTK * kDJ; e += 256) {
                    const int kk = e / kDJ, jj = e % kDJ;
                    const int k = kt * kTK + kk, j = jc * kDJ + jj;
                    es[kk * kES + jj] = (k < K && j < d) ? IO<T>::cb(__ldg(E + (size_t)k * d + j)) : 0.0f;
                }
            }
            __syncthreads();
            if (tid < kTK) {  // |e_k|^2 carried across chunks: same sequential order as the contract
#pragma unroll 8
                for (int jj = 0; jj < kDJ; ++jj) { const float v = es[tid * kES + jj]; ee_acc = fmaf(v, v, ee_acc); }
            }
            const float* zrow = zs + (size_t)(jc * kDJ) * TM + tm * 4;
#pragma unroll 2
            for (int jj = 0; jj < kDJ; jj += 4) {
                float4 e4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) e4[i] = *reinterpret_cast<const float4*>(es + (tk + 16 * i) * kES + jj);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float zr[RM];
                    const float4 za = *reinterpret_cast<const float4*>(zrow + (size_t)(jj + u) * TM);
                    zr[0] = za.x; zr[1] = za.y; zr[2] = za.z; zr[3] = za.w;
                    if (RM == 8) {
                        const float4 zb = *reinterpret_cast<const float4*>(zrow + (size_t)(jj + u) * TM + TM / 2);
                        zr[RM - 4] = zb.x; zr[RM - 3] = zb.y; zr[RM - 2] = zb.z; zr[RM - 1] = zb.w;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float ev = u == 0 ? e4[i].x : u == 1 ? e4[i].y : u == 2 ? e4[i].z : e4[i].w;
#pragma unroll
                        for (int r = 0; r < RM; ++r) acc[r][i] = fmaf(zr[r], ev, acc[r][i]);
                    }
                }
            }
        }
        if (tid < kTK) ee_s[tid] = (kt * kTK + tid < K) ? ee_acc : CUDART_INF_F;
        __syncthreads();
        // ---- epilogue of this code tile: thread-local ascending-k scan, then half-warp (16 lanes share a row)
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            const int m = (RM == 8) ? tm * 4 + (r & 3) + (r >> 2) * (TM / 2) : tm * 4 + r;
            const float zz = zz_s[m];
            float bv = CUDART_INF_F, sv = CUDART_INF_F;
            int bi = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = kt * kTK + tk + 16 * i;
                const float dist = dist_f32(zz, ee_s[tk + 16 * i], acc[r][i]);
                const bool take = (k < K) && (bi == 0x7fffffff || (!(dist >= bv) && (bv == bv)));  // the first code seeds the scan (all-+inf row -> lowest k)
                if (take) { sv = bv; bv = dist; bi = k; }  // the old best (<= old second) becomes the second
                else if (k < K) sv = fminf(sv, dist);
            }
            // merging two disjoint sets: second = min(second_a, second_b, max(best_a, best_b))
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const float os = __shfl_xor_sync(0xffffffffu, sv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                sv = fminf(fminf(sv, os), fmaxf(bv, ov));
                if (lex_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            secv[r] = fminf(fminf(secv[r], sv), kt == 0 ? CUDART_INF_F : fmaxf(bestv[r], bv));
            if (kt == 0 || lex_better(bv, bi, bestv[r], besti[r])) { bestv[r] = bv; besti[r] = bi; }
        }
    }
    if (tk == 0) {
        unsigned nnear = 0u;
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            const int m = (RM == 8) ? tm * 4 + (r & 3) + (r >> 2) * (TM / 2) : tm * 4 + r;
            idx_s[m] = besti[r];
            if (row0 + m < p.N && bestv[r] == bestv[r]) nnear += near_tie(bestv[r], secv[r]) ? 1u : 0u;
        }
        if (p.neartie && nnear) atomicAdd(p.neartie, (unsigned long long)nnear);  // rare: <~1 % of the rows
    }
    __syncthreads();
    if (tid < TM && lvalid) p.idx[seg][((size_t)lb * p.C + c) * HW + lp] = (long long)idx_s[tid];
    if (!p.fused) return;

    // ---- fused gather + straight-through + loss: z still resident in shared memory -----------------------
    float lsum = 0.0f;
    if (lvalid) {
        const int k = idx_s[lm];
        const float* e = E + (size_t)k * d;
        T* out = reinterpret_cast<T*>(p.q) + ((size_t)lb * p.C * d + (size_t)c * d) * HW + lp;
        constexpr int PARTS = 256 / TM;
        const int part = tid / TM;
        if (e_vec) {
            for (int j = part * 4; j < d; j += PARTS * 4) {
                const float4 q4 = __ldg(reinterpret_cast<const float4*>(e + j));
                const float qv[4] = {IO<T>::cb(q4.x), IO<T>::cb(q4.y), IO<T>::cb(q4.z), IO<T>::cb(q4.w)};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float zv = zs[(size_t)(j + u) * TM + lm];
                    const float diff = __fsub_rn(qv[u], zv);
                    IO<T>::st(out + (size_t)(j + u) * HW, __fadd_rn(zv, diff));  // z + (q - z), models/vq_vae.py:53
                    lsum = fmaf(diff, diff, lsum);
                }
            }
        } else {
            for (int j = part; j < d; j += PARTS) {
                const float zv = zs[(size_t)j * TM + lm];
                const float diff = __fsub_rn(IO<T>::cb(__ldg(e + j)), zv);
                IO<T>::st(out + (size_t)j * HW, __fadd_rn(zv, diff));
                lsum = fmaf(diff, diff, lsum);
            }
        }
    }
    block_accumulate((double)lsum, &p.loss_acc[c], red);
    if (last_block(p.ticket, gridDim.x * gridDim.y)) finalize_losses(p);
}

template <typename T>
static int launch_forward_simt_t(const QuantParams& p, cudaStream_t s) {
    const int dpad = (p.d + kDJ - 1) / kDJ * kDJ;
    auto smem_for = [&](int TM) { return (size_t)((size_t)dpad * TM + kTK * kES + TM + kTK + TM) * sizeof(float); };
    QuantParams q = p;
    cudaError_t e;
    if (smem_for(128) <= 200 * 1024) {
        q.tiles_per_seg = (int)((p.N + 127) / 128);
        const size_t sm = smem_for(128);
        e = cudaFuncSetAttribute(vq_fwd_simt_kernel<128, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((unsigned)(q.tiles_per_seg * p.n_seg), (unsigned)p.C);
        vq_fwd_simt_kernel<128, T><<<grid, 256, sm, s>>>(q);
    } else if (smem_for(64) <= 200 * 1024) {
        q.tiles_per_seg = (int)((p.N + 63) / 64);
        const size_t sm = smem_for(64);
        e = cudaFuncSetAttribute(vq_fwd_simt_kernel<64, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((unsigned)(q.tiles_per_seg * p.n_seg), (unsigned)p.C);
        vq_fwd_simt_kernel<64, T><<<grid, 256, sm, s>>>(q);
    } else {
        return CTVQ_E_UNSUPPORTED;
    }
    return (int)cudaGetLastError();
}

int launch_forward_simt(const QuantParams& p, cudaStream_t s) {
    return p.dtype == CTVQ_BF16 ? launch_forward_simt_t<__nv_bfloat16>(p, s) : launch_forward_simt_t<float>(p, s);
}

// ------------------------------------------------------------------------------------------------
// gather by supplied indices + straight-through + loss  (compute_latents)
// ------------------------------------------------------------------------------------------------
template <int VEC, typename T>
__global__ void __launch_bounds__(256) gather_st_loss_kernel(const QuantParams p) {
    __shared__ double red[32];
    const int c = blockIdx.y;
    const int d = p.d, HW = p.HW, HWV = HW / VEC;
    const T* __restrict__ z = reinterpret_cast<const T*>(p.z[0]);
    T* __restrict__ qout = reinterpret_cast<T*>(p.q);
    const float* __restrict__ E = p.E[c];
    const long long* __restrict__ idx = p.idx[0];
    const long long total = p.B * (long long)d * HWV;  // items of this codebook
    float lsum = 0.0f;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total;
         it += (long long)gridDim.x * blockDim.x) {
        const int pv = (int)(it % HWV);
        const long long bj = it / HWV;
        const int j = (int)(bj % d);
        const long long b = bj / d;
        const size_t zoff = ((size_t)b * p.Dtot + (size_t)c * p.cs + j) * HW + (size_t)pv * VEC;
        const size_t ooff = ((size_t)b * p.C * d + (size_t)c * d + j) * HW + (size_t)pv * VEC;
        const size_t ioff = ((size_t)b * p.C + c) * HW + (size_t)pv * VEC;
        float zv[4], ov[4];
        long long kv[VEC];
        if (VEC == 4) {
            IO<T>::ld4(z + zoff, zv);
            const longlong2 i0 = __ldg(reinterpret_cast<const longlong2*>(idx + ioff));
            const longlong2 i1 = __ldg(reinterpret_cast<const longlong2*>(idx + ioff + 2));
            kv[0] = i0.x; kv[1 % VEC] = i0.y; kv[2 % VEC] = i1.x; kv[3 % VEC] = i1.y;
        } else {
            zv[0] = IO<T>::ld(z + zoff);
            kv[0] = __ldg(idx + ioff);
        }
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            long long k = kv[u];
            if (k < 0 || k >= p.K) { atomicOr(p.err, 1u); k = k < 0 ? 0 : p.K - 1; }
            const float diff = __fsub_rn(IO<T>::cb(__ldg(E + (size_t)k * d + j)), zv[u]);
            ov[u] = __fadd_rn(zv[u], diff);
            lsum = fmaf(diff, diff, lsum);
        }
        if (VEC == 4) IO<T>::st4(qout + ooff, ov);
        else IO<T>::st(qout + ooff, ov[0]);
    }
    block_accumulate((double)lsum, &p.loss_acc[c], red);
    if (last_block(p.ticket, gridDim.x * gridDim.y)) finalize_losses(p);
}

template <typename T>
static int launch_gather_t(const QuantParams& p, cudaStream_t s) {
    const bool vec = (p.HW % 4 == 0) && IO<T>::aligned4(p.z[0]) && IO<T>::aligned4(p.q) && ((reinterpret_cast<uintptr_t>(p.idx[0]) & 15) == 0);
    const long long items = p.B * (long long)p.d * (p.HW / (vec ? 4 : 1));
    long long blocks = (items + 256 * 4 - 1) / (256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    dim3 grid((unsigned)blocks, (unsigned)p.C);
    if (vec) gather_st_loss_kernel<4, T><<<grid, 256, 0, s>>>(p);
    else gather_st_loss_kernel<1, T><<<grid, 256, 0, s>>>(p);
    return (int)cudaGetLastError();
}

int launch_gather(const QuantParams& p, cudaStream_t s) {
    return p.dtype == CTVQ_BF16 ? launch_gather_t<__nv_bfloat16>(p, s) : launch_gather_t<float>(p, s);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int VEC, typename T>
__global__ void __launch_bounds__(256) backward_kernel(const BwdParams p) {
    const T* __restrict__ zT = reinterpret_cast<const T*>(p.z);
    const T* __restrict__ goT = reinterpret_cast<const T*>(p.g_out);
    T* __restrict__ gzT = reinterpret_cast<T*>(p.gz);
    extern __shared__ float acc_s[];  // [C*K*d] when smem_acc
    const int d = p.d, HW = p.HW, HWV = HW / VEC, C = p.C, K = p.K;
    const int ckd = C * K * d;
    if (p.smem_acc) {
        for (int i = threadIdx.x; i < ckd; i += blockDim.x) acc_s[i] = 0.0f;
        __syncthreads();
    }
    const float gl = __ldg(p.g_loss);
    const double nd = (double)p.N * (double)d;
    const float coef_e = (float)(2.0 / nd) * gl;               // d vq_loss / d E   weight on (q - z)
    const float coef_z = (float)(2.0 * (double)p.beta / nd) * gl;  // d vq_loss / d z   weight on (z - q)
    const long long total = p.B * (long long)p.Dtot * HWV;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total;
         it += (long long)gridDim.x * blockDim.x) {
        const int pv = (int)(it % HWV);
        const long long bch = it / HWV;
        const int ch = (int)(bch % p.Dtot);
        const long long b = bch / p.Dtot;
        const size_t zoff = ((size_t)b * p.Dtot + ch) * HW + (size_t)pv * VEC;
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = 0.0f;
        // codebooks whose slice [c*cs, c*cs+d) contains ch
        int c_hi = p.cs > 0 ? ch / p.cs : 0;
        if (c_hi > C - 1) c_hi = C - 1;
        int c_lo = 0;
        if (ch - d + 1 > 0) c_lo = p.cs > 0 ? (ch - d + 1 + p.cs - 1) / p.cs : 0;
        if (p.cs == 0) { c_lo = 0; c_hi = (ch < d) ? C - 1 : -1; }
        if (c_lo <= c_hi) {
            float zv[4];
            if (VEC == 4) IO<T>::ld4(zT + zoff, zv);
            else zv[0] = IO<T>::ld(zT + zoff);
            for (int c = c_lo; c <= c_hi; ++c) {
                const int j = ch - c * p.cs;
                const size_t ioff = ((size_t)b * C + c) * HW + (size_t)pv * VEC;
                const size_t goff = ((size_t)b * C * d + (size_t)c * d + j) * HW + (size_t)pv * VEC;
                long long kv[VEC];
                float go[4] = {0.f, 0.f, 0.f, 0.f};
                if (VEC == 4) {
                    const longlong2 i0 = __ldg(reinterpret_cast<const longlong2*>(p.idx + ioff));
                    const longlong2 i1 = __ldg(reinterpret_cast<const longlong2*>(p.idx + ioff + 2));
                    kv[0] = i0.x; kv[1 % VEC] = i0.y; kv[2 % VEC] = i1.x; kv[3 % VEC] = i1.y;
                    if (goT) IO<T>::ld4(goT + goff, go);
                } else {
                    kv[0] = __ldg(p.idx + ioff);
                    go[0] = goT ? IO<T>::ld(goT + goff) : 0.0f;
                }
                const float* __restrict__ E = p.E[c];
#pragma unroll
                for (int u = 0; u < VEC; ++u) {
                    long long k = kv[u];
                    if (k < 0 || k >= K) { atomicOr(p.err, 1u); k = k < 0 ? 0 : K - 1; }
                    const float diff = __fsub_rn(IO<T>::cb(__ldg(E + (size_t)k * d + j)), zv[u]);  // q - z
                    g[u] += go[u] - coef_z * diff;
                    const int a = (c * K + (int)k) * d + j;
                    if (p.smem_acc) atomicAdd(&acc_s[a], diff);
                    else atomicAdd(&p.gE[a], coef_e * diff);
                }
            }
        }
        if (VEC == 4) IO<T>::st4(gzT + zoff, g);
        else IO<T>::st(gzT + zoff, g[0]);
    }
    if (p.smem_acc) {
        __syncthreads();
        for (int i = threadIdx.x; i < ckd; i += blockDim.x) {
            const float v = acc_s[i];
            if (v != 0.0f) atomicAdd(&p.gE[i], coef_e * v);
        }
    }
    peer_tail(p.peer, p.gE);  // fused collective (no-op unless ctvq_backward_allreduce armed it)
}

template <typename T>
static int launch_backward_direct(BwdParams p, cudaStream_t s) {
    const bool vec = (p.HW % 4 == 0) && IO<T>::aligned4(p.z) && IO<T>::aligned4(p.gz) && ((reinterpret_cast<uintptr_t>(p.idx) & 15) == 0) &&
                     (p.g_out == nullptr || IO<T>::aligned4(p.g_out));
    const long long items = p.B * (long long)p.Dtot * (p.HW / (vec ? 4 : 1));
    const size_t acc_bytes = sizeof(float) * (size_t)p.C * p.K * p.d;
    p.smem_acc = (acc_bytes <= 64 * 1024 && p.N >= 32768) ? 1 : 0;
    long long blocks = (items + 256 * 4 - 1) / (256 * 4);
    if (blocks < 1) blocks = 1;
    const long long cap = p.smem_acc ? sm_count() * 2 : sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const size_t sm = p.smem_acc ? acc_bytes : 0;
    cudaError_t e;
    if (vec) {
        if (sm > 48 * 1024) {
            e = cudaFuncSetAttribute(backward_kernel<4, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return (int)e;
        }
        backward_kernel<4, T><<<(unsigned)blocks, 256, sm, s>>>(p);
    } else {
        if (sm > 48 * 1024) {
            e = cudaFuncSetAttribute(backward_kernel<1, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return (int)e;
        }
        backward_kernel<1, T><<<(unsigned)blocks, 256, sm, s>>>(p);
    }
    return (int)cudaGetLastError();
}

int launch_backward(const BwdParams& p0, cudaStream_t s) {
    BwdParams p = p0;
    cudaError_t e = cudaMemsetAsync(p.gE, 0, sizeof(float) * (size_t)p.C * p.K * p.d, s);
    if (e != cudaSuccess) return (int)e;
    {
        int rc = CTVQ_E_UNSUPPORTED;
        if (p.dtype == CTVQ_F32) rc = launch_backward_ring(p, s);  // one codebook of many codes, large batch: resident accumulator + rings
        if (rc != CTVQ_E_UNSUPPORTED) return rc;
        if (p.dtype == CTVQ_F32) rc = launch_backward_c1(p, s);  // one full-width codebook at a batch that amortises a per-CTA accumulator
        if (rc != CTVQ_E_UNSUPPORTED) return rc;
        rc = launch_backward_fast(p, s);      // shape-specialised (configs' shapes), fp32 and bf16
        if (rc != CTVQ_E_UNSUPPORTED) return rc;
        rc = launch_backward_tiled(p, s);     // shared-memory accumulator, no atomics in the inner loop
        if (rc != CTVQ_E_UNSUPPORTED) return rc;
    }
    return p.dtype == CTVQ_BF16 ? launch_backward_direct<__nv_bfloat16>(p, s) : launch_backward_direct<float>(p, s);
}

// ------------------------------------------------------------------------------------------------
// Gaussian branch: reparameterise + KL
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reparam_kld_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                              const float* __restrict__ eps, long long n, long long B,
                                                              float* __restrict__ z, float* kld_out, double* acc,
                                                              unsigned int* ticket, int vec) {
    __shared__ double red[32];
    float part = 0.0f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const long long n4 = n >> 2;
        for (long long i = t0; i < n4; i += stride) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mu) + i);
            const float4 l = __ldg(reinterpret_cast<const float4*>(lv) + i);
            const float4 e = __ldg(reinterpret_cast<const float4*>(eps) + i);
            float4 o;
            o.x = fmaf(e.x, expf(0.5f * l.x), m.x);
            o.y = fmaf(e.y, expf(0.5f * l.y), m.y);
            o.z = fmaf(e.z, expf(0.5f * l.z), m.z);
            o.w = fmaf(e.w, expf(0.5f * l.w), m.w);
            reinterpret_cast<float4*>(z)[i] = o;
            part += (1.0f + l.x - m.x * m.x - expf(l.x)) + (1.0f + l.y - m.y * m.y - expf(l.y)) +
                    (1.0f + l.z - m.z * m.z - expf(l.z)) + (1.0f + l.w - m.w * m.w - expf(l.w));
        }
    } else {
        for (long long i = t0; i < n; i += stride) {
            const float m = __ldg(mu + i), l = __ldg(lv + i);
            z[i] = fmaf(__ldg(eps + i), expf(0.5f * l), m);
            part += 1.0f + l - m * m - expf(l);
        }
    }
    block_accumulate((double)part, acc, red);
    if (last_block(ticket, gridDim.x)) {
        __threadfence();
        if (threadIdx.x == 0) {
            *kld_out = (float)(-0.5 * __ldcg(acc) / (double)B);
            *acc = 0.0;
            *ticket = 0u;
            __threadfence();
        }
    }
}

__global__ void __launch_bounds__(256) reparam_kld_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                              const float* __restrict__ eps, const float* __restrict__ g_z,
                                                              const float* __restrict__ g_kld, long long n, long long B,
                                                              float* __restrict__ g_mu, float* __restrict__ g_lv) {
    const float gk = g_kld ? __ldg(g_kld) / (float)B : 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = __ldg(mu + i), l = __ldg(lv + i);
        const float gz = g_z ? __ldg(g_z + i) : 0.0f;
        g_mu[i] = fmaf(gk, m, gz);
        // d z / d lv = eps * 0.5 * exp(0.5 lv) ;  d kld / d lv = -0.5 (1 - exp(lv)) / B
        g_lv[i] = gz * __ldg(eps + i) * 0.5f * expf(0.5f * l) - 0.5f * gk * (1.0f - expf(l));
    }
}

int launch_reparam_fwd(const float* mu, const float* lv, const float* eps, long long B, int L, float* z, float* kld,
                       Workspace* ws, cudaStream_t s) {
    const long long n = B * (long long)L;
    const int vec = ((n & 3) == 0) && (((reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(lv) |
                                         reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(z)) & 15) == 0);
    long long blocks = ((vec ? n / 4 : n) + 256 * 4 - 1) / (256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    reparam_kld_fwd_kernel<<<(unsigned)blocks, 256, 0, s>>>(mu, lv, eps, n, B, z, kld, &ws->kld_acc, &ws->ticket2, vec);
    return (int)cudaGetLastError();
}

int launch_reparam_bwd(const float* mu, const float* lv, const float* eps, const float* g_z, const float* g_kld,
                       long long B, int L, float* g_mu, float* g_lv, cudaStream_t s) {
    const long long n = B * (long long)L;
    long long blocks = (n + 256 * 4 - 1) / (256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    reparam_kld_bwd_kernel<<<(unsigned)blocks, 256, 0, s>>>(mu, lv, eps, g_z, g_kld, n, B, g_mu, g_lv);
    return (int)cudaGetLastError();
}

}  // namespace ctvq
