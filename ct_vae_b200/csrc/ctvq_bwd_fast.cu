// Shape-specialised backward kernel (compile-time D, C, K, HW, Dtot, chan_stride) for the configs' shapes.
//
// Same maths as ctvq_bwd.cu (autograd of models/vq_vae.py:43-53 with the overlapping slices of
// models/mcq_vae.py:117), re-organised so the two halves of the work run CONCURRENTLY on different warps of a
// CTA once the tile's indices and z channels are staged in shared memory:
//   warps 4-7  grad_z: one warp per input channel, lanes along H*W (4 rows per lane): 128-bit loads of g_out,
//              128-bit stores of grad_z (channels no slice reads get zeros), codeword values from the padded
//              shared-memory copy of the codebooks;
//   warps 0-3  codebook gradient: each warp owns one (codebook, 32-channel chunk) slab of the shared [C,K,d]
//              accumulator, lanes along the channel — plain shared-memory read-modify-write, no atomics.
// All address arithmetic folds into immediates; the accumulator is flushed once per (persistent) CTA.
#include <stdlib.h>

#include "ctvq_tc_ptx.cuh"

namespace ctvq {
namespace {

constexpr int kBT = 256;
constexpr int kTM = 128;

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Warp-specialised: warps 0-3 ("acc") stage each tile's indices + touched z channels into a double-buffered
// shared-memory slot (register prefetch one tile ahead) and run the codebook-gradient accumulation; warps 4-7 ("gz")
// run grad_z with a 2-deep software pipeline of 128-bit g_out loads.  The two halves only meet at named barriers
// FULL[buf] / EMPTY[buf], so neither waits for the other inside a tile.
template <int D, int C, int K, int HWT, int DTOT, int CS, int MINB>
__global__ void __launch_bounds__(kBT, MINB) vq_bwd_fast_kernel(const BwdParams p, const int ntiles) {
    constexpr int USED = (C - 1) * CS + D;
    constexpr int ZS = kTM + 1;
    constexpr int ESD = D + 1;
    constexpr int CKD = C * K * D;
    constexpr int JCH = (D + 31) / 32;
    constexpr int ITEMS = C * JCH;
    constexpr int kFull = 1, kEmpty = 3, kAcc = 5;  // named barrier ids (0 = __syncthreads)
    static_assert(ITEMS <= 4, "one accumulation warp per (codebook, channel chunk)");
    static_assert(HWT % 4 == 0 && kTM % 4 == 0, "row quads");
    extern __shared__ __align__(16) float smem[];
    int* idx_s = reinterpret_cast<int*>(smem);      // [2][C][TM]
    float* zs = smem + 2 * C * kTM;                 // [2][USED][ZS]
    float* acc = zs + 2 * USED * ZS;                // [C][K][D]
    float* es = acc + CKD;                          // [C][K][D+1]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < CKD; i += kBT) acc[i] = 0.0f;
    for (int i = tid; i < CKD; i += kBT) {
        const int j = i % D, ck = i / D;
        es[ck * ESD + j] = __ldg(p.E[ck / K] + (size_t)(ck % K) * D + j);
    }
    const float gl = __ldg(p.g_loss);
    const double nd = (double)p.N * (double)D;
    const float coef_e = (float)(2.0 / nd) * gl;
    const float coef_z = (float)(2.0 * (double)p.beta / nd) * gl;
    __syncthreads();
    const int niter = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < 4) {
        // =========================== acc warps: staging + codebook-gradient accumulation ===========================
        float zreg[USED];
        long long kreg[C];
        const int m = tid;  // 0..127: this thread stages row m of every tile
        auto prefetch = [&](int it) {
            const long long n = (long long)(blockIdx.x + it * gridDim.x) * kTM + m;
            const bool valid = n < p.N;
            const long long b = valid ? n / HWT : 0;
            const int hw = valid ? (int)(n - b * HWT) : 0;
#pragma unroll
            for (int c = 0; c < C; ++c) kreg[c] = valid ? __ldg(p.idx + ((size_t)b * C + c) * HWT + hw) : 0;
            const float* src = p.z + (size_t)b * DTOT * HWT + hw;
#pragma unroll
            for (int ch = 0; ch < USED; ++ch) zreg[ch] = valid ? __ldg(src + (size_t)ch * HWT) : 0.0f;
        };
        if (niter > 0) prefetch(0);
        for (int it = 0; it < niter; ++it) {
            const int buf = it & 1;
            const long long row0 = (long long)(blockIdx.x + it * gridDim.x) * kTM;
            const int mcount = (int)min((long long)kTM, p.N - row0);
            if (it >= 2) named_sync(kEmpty + buf, kBT);  // gz warps finished reading this slot (iteration it-2)
            int* idb = idx_s + buf * C * kTM;
            float* zb = zs + buf * USED * ZS;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                long long kk = kreg[c];
                if (kk < 0 || kk >= K) { atomicOr(p.err, 1u); kk = kk < 0 ? 0 : K - 1; }
                idb[c * kTM + m] = (int)kk;
            }
#pragma unroll
            for (int ch = 0; ch < USED; ++ch) zb[ch * ZS + m] = zreg[ch];
            named_arrive(kFull + buf, kBT);  // slot ready for the gz warps
            named_sync(kAcc, 128);           // ... and for the other acc warps
            if (it + 1 < niter) prefetch(it + 1);  // latency hidden behind this tile's accumulation
            if (warp < ITEMS) {
                const int c = warp / JCH;
                const int j = (warp - c * JCH) * 32 + lane;
                const bool act = j < D;
                const int jj = act ? j : 0;
                const float* zcol = zb + (c * CS + jj) * ZS;
                const int* ks = idb + c * kTM;
                float* ac = acc + c * K * D + jj;
                const float* ec = es + c * K * ESD + jj;
                int r = 0;
                for (; r + 4 <= mcount; r += 4) {
                    const int4 kk = *reinterpret_cast<const int4*>(ks + r);
                    const bool distinct = kk.x != kk.y && kk.x != kk.z && kk.x != kk.w && kk.y != kk.z && kk.y != kk.w &&
                                          kk.z != kk.w;
                    if (act) {
                        const float d0 = __fsub_rn(ec[kk.x * ESD], zcol[r]), d1 = __fsub_rn(ec[kk.y * ESD], zcol[r + 1]);
                        const float d2 = __fsub_rn(ec[kk.z * ESD], zcol[r + 2]), d3 = __fsub_rn(ec[kk.w * ESD], zcol[r + 3]);
                        if (distinct) {
                            const float a0 = ac[kk.x * D], a1 = ac[kk.y * D], a2 = ac[kk.z * D], a3 = ac[kk.w * D];
                            ac[kk.x * D] = a0 + d0; ac[kk.y * D] = a1 + d1; ac[kk.z * D] = a2 + d2; ac[kk.w * D] = a3 + d3;
                        } else {
                            ac[kk.x * D] += d0; ac[kk.y * D] += d1; ac[kk.z * D] += d2; ac[kk.w * D] += d3;
                        }
                    }
                }
                for (; r < mcount; ++r)
                    if (act) ac[ks[r] * D] += __fsub_rn(ec[ks[r] * ESD], zcol[r]);
            }
        }
    } else {
        // =========================== gz warps: grad_z, lanes along H*W ==========================================
        const int gw = warp - 4;
        const int m = lane * 4;
        for (int it = 0; it < niter; ++it) {
            const int buf = it & 1;
            const long long row0 = (long long)(blockIdx.x + it * gridDim.x) * kTM;
            const int mcount = (int)min((long long)kTM, p.N - row0);
            const int* idb = idx_s + buf * C * kTM;
            const float* zb = zs + buf * USED * ZS;
            // the zero channels need nothing from shared memory: write them before waiting for the slot
            if (m < mcount) {
                const long long n = row0 + m;
                const long long b = n / HWT;
                float* gz_row = p.gz + (size_t)b * DTOT * HWT + (int)(n - b * HWT);
                for (int ch = USED + (gw - USED % 4 + 4) % 4; ch < DTOT; ch += 4)
                    *reinterpret_cast<float4*>(gz_row + (size_t)ch * HWT) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            named_sync(kFull + buf, kBT);
            if (m < mcount) {
                const long long n = row0 + m;
                const long long b = n / HWT;
                const int hw = (int)(n - b * HWT);
                const float* go_row = p.g_out ? p.g_out + (size_t)b * C * D * HWT + hw : nullptr;
                float* gz_row = p.gz + (size_t)b * DTOT * HWT + hw;
                int eb[C][4];  // shared-memory float index of this row's codeword per codebook
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int4 kk = *reinterpret_cast<const int4*>(idb + c * kTM + m);
                    eb[c][0] = (c * K + kk.x) * ESD; eb[c][1] = (c * K + kk.y) * ESD;
                    eb[c][2] = (c * K + kk.z) * ESD; eb[c][3] = (c * K + kk.w) * ESD;
                }
                auto load_go = [&](int ch, float4 (&go)[C]) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = ch - c * CS;
                        go[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ch < USED && j >= 0 && j < D && go_row)
                            go[c] = __ldg(reinterpret_cast<const float4*>(go_row + (size_t)(c * D + j) * HWT));
                    }
                };
                // active channels ch = gw, gw+4, ...: 2-deep software pipeline of the g_out loads
                float4 g0[C], g1[C], g2[C];
                int ch = gw;
                load_go(ch, g0);
                load_go(ch + 4, g1);
                for (; ch < USED; ch += 4) {
                    load_go(ch + 8, g2);
                    const float z0 = zb[ch * ZS + m], z1 = zb[ch * ZS + m + 1], z2 = zb[ch * ZS + m + 2], z3 = zb[ch * ZS + m + 3];
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = ch - c * CS;
                        if (j >= 0 && j < D) {
                            const float d0 = __fsub_rn(es[eb[c][0] + j], z0), d1 = __fsub_rn(es[eb[c][1] + j], z1);
                            const float d2 = __fsub_rn(es[eb[c][2] + j], z2), d3 = __fsub_rn(es[eb[c][3] + j], z3);
                            g.x += g0[c].x - coef_z * d0; g.y += g0[c].y - coef_z * d1;
                            g.z += g0[c].z - coef_z * d2; g.w += g0[c].w - coef_z * d3;
                        }
                    }
                    *reinterpret_cast<float4*>(gz_row + (size_t)ch * HWT) = g;
#pragma unroll
                    for (int c = 0; c < C; ++c) { g0[c] = g1[c]; g1[c] = g2[c]; }
                }
            }
            named_arrive(kEmpty + buf, kBT);  // slot may be overwritten
        }
    }
    __syncthreads();
    for (int i = tid; i < CKD; i += kBT) {
        const float v = acc[i];
        if (v != 0.0f) atomicAdd(&p.gE[i], coef_e * v);
    }
    peer_tail(p.peer, p.gE);  // fused collective (no-op unless ctvq_backward_allreduce armed it)
}

template <int D, int C, int K, int HWT, int DTOT, int CS, int MINB>
int launch(const BwdParams& p, cudaStream_t s) {
    constexpr int USED = (C - 1) * CS + D;
    constexpr size_t smem = sizeof(float) * (2 * (size_t)C * kTM + 2 * (size_t)USED * (kTM + 1) + (size_t)C * K * D + (size_t)C * K * (D + 1));
    static_assert(smem <= (MINB == 2 ? 113 : 225) * 1024, "shared memory budget");
    const long long nt = (p.N + kTM - 1) / kTM;
    if (nt > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
    int grid = sm_count() * MINB;
    if (grid > nt) grid = (int)nt;
    if (grid < 1) grid = 1;
    auto kern = vq_bwd_fast_kernel<D, C, K, HWT, DTOT, CS, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kBT, smem, s>>>(p, (int)nt);
    return (int)cudaGetLastError();
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {  // hinted wait, ~2 s bound then trap
    for (int it = 0; it < 2048; ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// four consecutive elements of a shared-memory row as floats (fp32: one LDS.128; bf16: one LDS.64, exact widening)
__device__ __forceinline__ void lds4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void lds4(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}

// One image (H*W = TM rows) per tile, one persistent CTA per SM, 12 warps:
//   warps 0-3  "acc": register-prefetch the tile's indices + touched z channels one tile ahead, publish them in a
//              double-buffered shared slot, accumulate the codebook gradient (warp-owned slabs, no atomics);
//              thread 0 also drives a 3-stage cp.async.bulk (TMA) ring that streams the image's whole g_out block
//              (C*D x HW fp32, contiguous in NCHW) into shared memory two tiles ahead;
//   warps 4-11 "gz": grad_z from shared memory only (128-bit LDS of g_out, codeword gathers), 128-bit stores.
// No global load sits on any warp's critical path, so HBM stays busy with ~2 tiles of reads in flight per SM.
constexpr int kBT4 = 384;
// T: element type of z / g_out / grad_z (float, or __nv_bfloat16 for dtype = CTVQ_BF16: half the streamed bytes; codebook
// values are rounded to bf16 as they are staged, all arithmetic stays fp32)
template <int D, int C, int K, int HWT, int DTOT, int CS, typename T>
__global__ void __launch_bounds__(kBT4, 1) vq_bwd_tma_kernel(const BwdParams p, const int ntiles, const __grid_constant__ CUtensorMap gomap) {
    const T* __restrict__ zT = reinterpret_cast<const T*>(p.z);
    const T* __restrict__ goT = reinterpret_cast<const T*>(p.g_out);
    T* __restrict__ gzT = reinterpret_cast<T*>(p.gz);
    constexpr int TM = 64;                  // rows per tile: one 64-position segment of an image (the whole image at H*W = 64)
    constexpr int SEG = HWT / TM;           // tiles per image
    constexpr int NST = 3;                  // g_out ring depth
    constexpr int USED = (C - 1) * CS + D;
    constexpr int ZS = TM + 1;
    constexpr int ESD = D + 1;
    constexpr int CKD = C * K * D;
    constexpr int JCH = (D + 31) / 32;
    constexpr int ITEMS = C * JCH;
    constexpr int GOF = C * D * TM;         // floats per g_out stage
    constexpr int NGZ = 4;                  // gz warps (8..11)
    constexpr int NACC = 8;                 // acc warps: 2 per accumulator slab (row halves, private copies)
    constexpr int kFull = 1, kEmpty = 3, kAcc = 5;
    static_assert(ITEMS <= 4 && HWT % TM == 0, "configs' shapes");
    extern __shared__ __align__(128) float smem[];
    T* go_s = reinterpret_cast<T*>(smem);                 // [NST][C*D][TM] in the I/O element type
    int* idx_s = reinterpret_cast<int*>(smem + NST * GOF * sizeof(T) / sizeof(float));  // [2][C][TM]
    float* zs = reinterpret_cast<float*>(idx_s + 2 * C * TM);  // [2][USED][ZS]
    float* acc = zs + 2 * USED * ZS;                      // [2][C][K][D]  (one copy per row half)
    float* es = acc + 2 * CKD;                            // [C][K][D+1]
    uint64_t* bars = reinterpret_cast<uint64_t*>(es + ((C * K * ESD + 1) & ~1));  // full[NST], empty[NST]
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[NST]);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, NGZ); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // codebook staging: every thread's 128-bit loads are issued BEFORE its first store (one L2 / HBM round trip for the whole
    // [C,K,D] block instead of one per unrolled group of scalar loads: the codebooks were evicted by the forward's 0.5 GB
    // write stream, so each dependent round trip costs ~0.6 us of every launch)
    bool e16 = (D % 4 == 0);
#pragma unroll
    for (int c = 0; c < C; ++c) e16 = e16 && ((reinterpret_cast<uintptr_t>(p.E[c]) & 15) == 0);
    if (e16) {
        constexpr int D4 = D / 4, NV = (CKD / 4 + kBT4 - 1) / kBT4;
        float4 v[NV];
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int i = tid + u * kBT4;
            if (i < CKD / 4) {
                const int ck = i / D4;
                v[u] = __ldg(reinterpret_cast<const float4*>(p.E[ck / K] + (size_t)(ck % K) * D) + (i - ck * D4));
            }
        }
        for (int i = tid; i < 2 * CKD; i += kBT4) acc[i] = 0.0f;  // (scalar: the accumulators are only 8-byte aligned behind zs)
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int i = tid + u * kBT4;
            if (i < CKD / 4) {
                const int ck = i / D4;
                float* dst = es + ck * ESD + 4 * (i - ck * D4);
                dst[0] = IO<T>::cb(v[u].x); dst[1] = IO<T>::cb(v[u].y); dst[2] = IO<T>::cb(v[u].z); dst[3] = IO<T>::cb(v[u].w);
            }
        }
    } else {
        for (int i = tid; i < 2 * CKD; i += kBT4) acc[i] = 0.0f;
        for (int i = tid; i < CKD; i += kBT4) {
            const int j = i % D, ck = i / D;
            es[ck * ESD + j] = IO<T>::cb(__ldg(p.E[ck / K] + (size_t)(ck % K) * D + j));
        }
    }
    const float gl = __ldg(p.g_loss);
    const double nd = (double)p.N * (double)D;
    const float coef_e = (float)(2.0 / nd) * gl;
    const float coef_z = (float)(2.0 * (double)p.beta / nd) * gl;
    __syncthreads();
    const int niter = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool has_go = p.g_out != nullptr;

    if (warp < NACC) {
        // =========================== acc warps ===========================
        constexpr int NZ = (USED + 3) / 4, NI = (C + 3) / 4;
        float zreg[NZ];
        long long kreg[NI];
        const int m = tid & (TM - 1), half = tid / TM;  // thread stages row m, channels / codebooks congruent to `half` mod 4
        auto prefetch = [&](int it) {
            const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;  // tile = (image, segment)
            const long long b = t / SEG;
            const int r0 = (int)(t - b * SEG) * TM;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int c = half + 4 * i;
                kreg[i] = (c < C) ? __ldg(p.idx + ((size_t)b * C + c) * HWT + r0 + m) : 0;
            }
            const T* src = zT + (size_t)b * DTOT * HWT + r0 + m;
#pragma unroll
            for (int i = 0; i < NZ; ++i) {
                const int ch = half + 4 * i;
                zreg[i] = (ch < USED) ? IO<T>::ld(src + (size_t)ch * HWT) : 0.0f;
            }
        };
        auto issue_go = [&](int it) {  // warp 0: stream tile it's g_out block into ring slot it % NST
            const int st = it % NST;
            if (it >= NST) mbar_wait(bar_empty + 8 * st, (uint32_t)((it / NST) - 1) & 1u);
            const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
            const long long b = t / SEG;
            const int r0 = (int)(t - b * SEG) * TM;
            if (lane == 0) mbar_expect_tx(bar_full + 8 * st, GOF * (uint32_t)sizeof(T));
            __syncwarp();
            if (SEG == 1) {  // the image's whole g_out block is contiguous in NCHW: one bulk copy
                if (lane == 0) bulk_g2s(smem_u32(go_s + st * GOF), goT + (size_t)b * GOF, GOF * (uint32_t)sizeof(T), bar_full + 8 * st);
            } else {         // one 64-position run per channel: ONE 3-D tensor-map box [C*D][64] (a bulk copy per channel measured
                             // ~46 cycles of TMA service each -- 3 us per tile at 128 channels)
                if (lane == 0) tc::tma_load_3d(smem_u32(go_s + st * GOF), &gomap, bar_full + 8 * st, r0, 0, (int)b);
            }
            __syncwarp();
        };
        if (niter > 0) prefetch(0);
        if (warp == 0 && has_go)
            for (int it = 0; it < NST - 1 && it < niter; ++it) issue_go(it);
        for (int it = 0; it < niter; ++it) {
            const int buf = it & 1;
            if (warp == 0 && has_go && it + NST - 1 < niter) issue_go(it + NST - 1);
            if (it >= 2) named_sync(kEmpty + buf, kBT4 - 0);  // gz warps finished reading this z/idx slot (iteration it-2)
            int* idb = idx_s + buf * C * TM;
            float* zb = zs + buf * USED * ZS;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int c = half + 4 * i;
                if (c < C) {
                    long long kk = kreg[i];
                    if (kk < 0 || kk >= K) { atomicOr(p.err, 1u); kk = kk < 0 ? 0 : K - 1; }
                    idb[c * TM + m] = (int)kk;
                }
            }
#pragma unroll
            for (int i = 0; i < NZ; ++i)
                if (half + 4 * i < USED) zb[(half + 4 * i) * ZS + m] = zreg[i];
            named_arrive(kFull + buf, kBT4);
            named_sync(kAcc, NACC * 32);
            if (it + 1 < niter) prefetch(it + 1);
            if ((warp & 3) < ITEMS) {
                const int item = warp & 3, rh = warp >> 2;  // accumulator slab, row half
                const int c = item / JCH;
                const int j = (item - c * JCH) * 32 + lane;
                const bool act = j < D;
                const int jj = act ? j : 0;
                const float* zcol = zb + (c * CS + jj) * ZS + rh * (TM / 2);
                const int* ks = idb + c * TM + rh * (TM / 2);
                float* ac = acc + rh * CKD + c * K * D + jj;
                const float* ec = es + c * K * ESD + jj;
                // software pipeline, two row-groups ahead: the (index -> codeword -> diff) loads of groups g+1, g+2 are
                // issued before the read-modify-write of group g, so shared-memory latency overlaps instead of chaining
                auto diffs = [&](const int4& kk, int r) {
                    return make_float4(__fsub_rn(ec[kk.x * ESD], zcol[r]), __fsub_rn(ec[kk.y * ESD], zcol[r + 1]),
                                       __fsub_rn(ec[kk.z * ESD], zcol[r + 2]), __fsub_rn(ec[kk.w * ESD], zcol[r + 3]));
                };
                constexpr int NG = TM / 8;  // this warp's half of the rows
                int4 kA = *reinterpret_cast<const int4*>(ks), kB = *reinterpret_cast<const int4*>(ks + 4);
                float4 dA = diffs(kA, 0), dB = diffs(kB, 4);
#pragma unroll 4
                for (int g = 0; g < NG; ++g) {
                    int4 kC = kB;
                    float4 dC = dB;
                    if (g + 2 < NG) {
                        kC = *reinterpret_cast<const int4*>(ks + (g + 2) * 4);
                        dC = diffs(kC, (g + 2) * 4);
                    }
                    if (act) {
                        const bool distinct = kA.x != kA.y && kA.x != kA.z && kA.x != kA.w && kA.y != kA.z && kA.y != kA.w &&
                                              kA.z != kA.w;
                        if (distinct) {
                            const float a0 = ac[kA.x * D], a1 = ac[kA.y * D], a2 = ac[kA.z * D], a3 = ac[kA.w * D];
                            ac[kA.x * D] = a0 + dA.x; ac[kA.y * D] = a1 + dA.y; ac[kA.z * D] = a2 + dA.z; ac[kA.w * D] = a3 + dA.w;
                        } else {
                            ac[kA.x * D] += dA.x; ac[kA.y * D] += dA.y; ac[kA.z * D] += dA.z; ac[kA.w * D] += dA.w;
                        }
                    }
                    kA = kB; dA = dB; kB = kC; dB = dC;
                }
            }
        }
    } else {
        // =========================== gz warps ===========================
        const int gw = warp - NACC;              // 0..3
        const int hsel = lane >> 4;              // half-warp: which channel of the pair
        const int m = (lane & 15) * 4;           // rows m..m+3
        for (int it = 0; it < niter; ++it) {
            const int buf = it & 1, st = it % NST;
            const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
            const long long b = t / SEG;
            const int* idb = idx_s + buf * C * TM;
            const float* zb = zs + buf * USED * ZS;
            const T* gos = go_s + st * GOF;
            T* gz_row = gzT + (size_t)b * DTOT * HWT + (int)(t - b * SEG) * TM + m;
            // zero channels first: they depend on nothing
            const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int ch = USED + 2 * gw + hsel; ch < DTOT; ch += 2 * NGZ) IO<T>::st4(gz_row + (size_t)ch * HWT, zero4);
            named_sync(kFull + buf, kBT4);
            if (has_go) mbar_wait(bar_full + 8 * st, (uint32_t)(it / NST) & 1u);
            int eb[C][4];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int4 kk = *reinterpret_cast<const int4*>(idb + c * TM + m);
                eb[c][0] = (c * K + kk.x) * ESD; eb[c][1] = (c * K + kk.y) * ESD;
                eb[c][2] = (c * K + kk.z) * ESD; eb[c][3] = (c * K + kk.w) * ESD;
            }
            for (int ch = 2 * gw + hsel; ch < USED; ch += 2 * NGZ) {
                const float z0 = zb[ch * ZS + m], z1 = zb[ch * ZS + m + 1], z2 = zb[ch * ZS + m + 2], z3 = zb[ch * ZS + m + 3];
                float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int j = ch - c * CS;
                    if (j >= 0 && j < D) {
                        float go[4] = {0.f, 0.f, 0.f, 0.f};
                        if (has_go) lds4(gos + (c * D + j) * TM + m, go);
                        const float d0 = __fsub_rn(es[eb[c][0] + j], z0), d1 = __fsub_rn(es[eb[c][1] + j], z1);
                        const float d2 = __fsub_rn(es[eb[c][2] + j], z2), d3 = __fsub_rn(es[eb[c][3] + j], z3);
                        g[0] += go[0] - coef_z * d0; g[1] += go[1] - coef_z * d1;
                        g[2] += go[2] - coef_z * d2; g[3] += go[3] - coef_z * d3;
                    }
                }
                IO<T>::st4(gz_row + (size_t)ch * HWT, g);
            }
            __syncwarp();
            if (lane == 0 && has_go) mbar_arrive(bar_empty + 8 * st);  // ring slot may be refilled
            named_arrive(kEmpty + buf, kBT4);
        }
    }
    __syncthreads();
    for (int i = tid; i < CKD; i += kBT4) {
        const float v = acc[i] + acc[CKD + i];
        if (v != 0.0f) atomicAdd(&p.gE[i], coef_e * v);
    }
    peer_tail(p.peer, p.gE);  // fused collective: the last CTA pushes the finished gradient to every peer and reduces
}

template <int D, int C, int K, int HWT, int DTOT, int CS, typename T>
int launch_tma(const BwdParams& p, cudaStream_t s) {
    constexpr int USED = (C - 1) * CS + D;
    constexpr int TM = 64;
    constexpr size_t smem = sizeof(T) * 3 * (size_t)C * D * TM +
                            sizeof(float) * (2 * (size_t)C * TM + 2 * (size_t)USED * (TM + 1) +
                                             2 * (size_t)C * K * D + (((size_t)C * K * (D + 1) + 1) & ~(size_t)1)) + 6 * 8;
    static_assert(smem <= 227 * 1024, "shared memory");
    if (p.N % HWT != 0) return CTVQ_E_UNSUPPORTED;
    const long long nt = p.N / TM;  // one 64-position segment of an image per tile
    if (nt > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
    int grid = sm_count();
    if (grid > nt) grid = (int)nt;
    CUtensorMap gomap;
    memset(&gomap, 0, sizeof(gomap));
    if (HWT != TM && p.g_out != nullptr) {
        if (p.B > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
        const int rc = tc::make_plain_map(gomap, p.g_out, IO<T>::kDtype, HWT, (long long)C * D, p.B, TM, C * D);
        if (rc != CTVQ_OK) return rc;
    }
    auto kern = vq_bwd_tma_kernel<D, C, K, HWT, DTOT, CS, T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kBT4, smem, s>>>(p, (int)nt, gomap);
    return (int)cudaGetLastError();
}

// Single full-width codebook with FEW codes and WIDE rows (configs/ct_mcq_vae.yaml: C=1, d=128, K=64, latents [B,128,8,8]).
// Same skeleton as vq_bwd_tma_kernel (one image per tile, persistent CTA, g_out streamed by a cp.async.bulk ring), with
// three changes that make a 128-channel tile fit and balance the two halves:
//   * z is staged by 4-byte cp.async copies straight into the padded [D][TM+1] tile (no register prefetch: 128 channels
//     per row would cost 64 registers per staging thread), TWO tiles ahead, double-buffered, issued by the gz warps (which
//     have the lighter half of the work) and completed on an mbarrier (cp.async.mbarrier.arrive.noinc);
//   * the JCH = D/32 "acc" warps (lane = channel, warp = 32-channel chunk of the shared [K,D] accumulator: race-free by
//     ownership, plain read-modify-write -- fp32 shared atomics compile to a compare-and-swap loop, ATOMS.CAST.SPIN, on
//     sm_100a) leave q - z IN PLACE of z, so the NGZ "gz" warps compute grad_z = g_out - coef_z (q - z) from shared memory only
//     -- no index decode, no codebook copy padded for row-wise gathers.
template <int D, int K, int HWT, int NST>
__global__ void __launch_bounds__(256, 1) vq_bwd_c1_tma_kernel(const BwdParams p, const int ntiles, const __grid_constant__ CUtensorMap gomap) {
    constexpr int TM = 64;                 // rows per tile: one 64-position segment of an image (the whole image at H*W = 64)
    constexpr int SEG = HWT / TM;          // tiles per image
    constexpr int ZS = TM + 1;
    constexpr int KD = K * D;
    constexpr int JCH = D / 32;            // acc warps
    constexpr int NGZ = 8 - JCH;           // gz warps
    constexpr int NAT = JCH * 32, NGT = NGZ * 32;
    constexpr int GOF = D * TM;            // floats per g_out stage
    constexpr int kDiff = 1, kAcc = 3, kGz = 4;  // named barrier ids (kDiff + buf)
    static_assert(D % 32 == 0 && JCH >= 1 && JCH <= 4 && HWT % TM == 0 && (D * TM) % NGT == 0, "shape");
    extern __shared__ __align__(128) float smem[];
    float* go_s = smem;                                        // [NST][D][TM]
    int* idx_s = reinterpret_cast<int*>(go_s + NST * GOF);     // [2][TM]
    float* zs = reinterpret_cast<float*>(idx_s + 2 * TM);      // [2][D][ZS]: z, then q - z in place
    float* acc = zs + 2 * D * ZS;                              // [K][D]
    float* es = acc + KD;                                      // [K][D]  (read with lanes along the channel only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(es + KD);     // full[NST], empty[NST], zfull[2]
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[NST]), bar_zfull = smem_u32(&bars[2 * NST]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, NGZ); }
        for (int i = 0; i < 2; ++i) mbar_init(bar_zfull + 8 * i, NGT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < KD; i += 256) { acc[i] = 0.0f; es[i] = __ldg(p.E[0] + i); }
    const float gl = __ldg(p.g_loss);
    const double nd = (double)p.N * (double)D;
    const float coef_e = (float)(2.0 / nd) * gl;
    const float coef_z = (float)(2.0 * (double)p.beta / nd) * gl;
    __syncthreads();
    const int niter = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool has_go = p.g_out != nullptr;

    if (warp < JCH) {
        // =========================== acc warps: codebook-gradient accumulation, q - z left in place ===========================
        auto issue_go = [&](int it) {  // thread 0: stream tile it's g_out block into ring slot it % NST
            const int st = it % NST;
            if (it >= NST) mbar_wait(bar_empty + 8 * st, (uint32_t)((it / NST) - 1) & 1u);
            const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
            const long long b = t / SEG;
            mbar_expect_tx(bar_full + 8 * st, GOF * 4u);
            if (SEG == 1) bulk_g2s(smem_u32(go_s + st * GOF), p.g_out + (size_t)b * GOF, GOF * 4u, bar_full + 8 * st);
            else tc::tma_load_3d(smem_u32(go_s + st * GOF), &gomap, bar_full + 8 * st, (int)(t - b * SEG) * TM, 0, (int)b);  // box [D][64]
        };
        if (tid == 0 && has_go)
            for (int it = 0; it < NST - 1 && it < niter; ++it) issue_go(it);
        for (int it = 0; it < niter; ++it) {
            const int buf = it & 1;
            if (tid == 0 && has_go && it + NST - 1 < niter) issue_go(it + NST - 1);
            if (tid < TM) {  // this tile's indices (range-checked like every consumer of caller-supplied indices)
                const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
                const long long b = t / SEG;
                long long kk = __ldg(p.idx + (size_t)b * HWT + (int)(t - b * SEG) * TM + tid);
                if (kk < 0 || kk >= K) { atomicOr(p.err, 1u); kk = kk < 0 ? 0 : K - 1; }
                idx_s[buf * TM + tid] = (int)kk;
            }
            mbar_wait(bar_zfull + 8 * buf, (uint32_t)(it >> 1) & 1u);  // the gz warps' cp.async copies of this tile have landed
            named_sync(kAcc, NAT);                                      // ... and the indices are published
            float* zb = zs + buf * D * ZS + (warp * 32 + lane) * ZS;    // this lane's channel column
            const int* ks = idx_s + buf * TM;
            float* ac = acc + warp * 32 + lane;
            const float* ec = es + warp * 32 + lane;
            auto diffs = [&](const int4& kk, int r) {
                return make_float4(__fsub_rn(ec[kk.x * D], zb[r]), __fsub_rn(ec[kk.y * D], zb[r + 1]),
                                   __fsub_rn(ec[kk.z * D], zb[r + 2]), __fsub_rn(ec[kk.w * D], zb[r + 3]));
            };
            // software pipeline, two row-groups ahead (as in vq_bwd_tma_kernel): loads of groups g+1, g+2 fly under the
            // read-modify-write of group g; four rows update four DIFFERENT words unless two of them chose the same code
            constexpr int NG = TM / 4;
            int4 kA = *reinterpret_cast<const int4*>(ks), kB = *reinterpret_cast<const int4*>(ks + 4);
            float4 dA = diffs(kA, 0), dB = diffs(kB, 4);
#pragma unroll 4
            for (int g = 0; g < NG; ++g) {
                int4 kC = kB;
                float4 dC = dB;
                if (g + 2 < NG) {
                    kC = *reinterpret_cast<const int4*>(ks + (g + 2) * 4);
                    dC = diffs(kC, (g + 2) * 4);
                }
                const bool distinct = kA.x != kA.y && kA.x != kA.z && kA.x != kA.w && kA.y != kA.z && kA.y != kA.w && kA.z != kA.w;
                if (distinct) {
                    const float a0 = ac[kA.x * D], a1 = ac[kA.y * D], a2 = ac[kA.z * D], a3 = ac[kA.w * D];
                    ac[kA.x * D] = a0 + dA.x; ac[kA.y * D] = a1 + dA.y; ac[kA.z * D] = a2 + dA.z; ac[kA.w * D] = a3 + dA.w;
                } else {
                    ac[kA.x * D] += dA.x; ac[kA.y * D] += dA.y; ac[kA.z * D] += dA.z; ac[kA.w * D] += dA.w;
                }
                zb[g * 4] = dA.x; zb[g * 4 + 1] = dA.y; zb[g * 4 + 2] = dA.z; zb[g * 4 + 3] = dA.w;  // q - z for the gz warps
                kA = kB; dA = dB; kB = kC; dB = dC;
            }
            named_arrive(kDiff + buf, 256);                // this tile's differences are in place
        }
    } else {
        // =========================== gz warps: z staging two tiles ahead; grad_z = g_out - coef_z (q - z) ========================
        const int gw = warp - JCH, gt = tid - NAT;
        const int hsel = lane >> 4;              // half-warp: which channel of the pair
        const int m = (lane & 15) * 4;           // rows m..m+3
        constexpr int CPT = D * TM / NGT;        // 4-byte copies per thread per tile
        auto stage = [&](int it) {               // tile it -> buffer it & 1 (asynchronous; completion arrives on zfull[buf])
            const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
            const long long b = t / SEG;
            const int buf = it & 1;
            const float* src = p.z + (size_t)b * D * HWT + (int)(t - b * SEG) * TM;
            float* dst = zs + buf * D * ZS;
#pragma unroll 8
            for (int i = 0; i < CPT; ++i) {
                const int e = i * NGT + gt;      // element of the [D][TM] block: coalesced along H*W
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + (e / TM) * ZS + (e % TM))),
                             "l"(src + (size_t)(e / TM) * HWT + (e % TM)) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_zfull + 8 * buf) : "memory");
        };
        for (int it = 0; it < 2 && it < niter; ++it) stage(it);
        for (int it = 0; it < niter; ++it) {
            const int buf = it & 1, st = it % NST;
            const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
            const long long b = t / SEG;
            const float* zb = zs + buf * D * ZS;
            const float* gos = go_s + st * GOF;
            float* gz_row = p.gz + (size_t)b * D * HWT + (int)(t - b * SEG) * TM + m;
            named_sync(kDiff + buf, 256);
            if (has_go) mbar_wait(bar_full + 8 * st, (uint32_t)(it / NST) & 1u);
#pragma unroll 4
            for (int ch = 2 * gw + hsel; ch < D; ch += 2 * NGZ) {
                const float d0 = zb[ch * ZS + m], d1 = zb[ch * ZS + m + 1], d2 = zb[ch * ZS + m + 2], d3 = zb[ch * ZS + m + 3];
                float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_go) go = *reinterpret_cast<const float4*>(gos + ch * TM + m);
                *reinterpret_cast<float4*>(gz_row + (size_t)ch * HWT) =
                    make_float4(go.x - coef_z * d0, go.y - coef_z * d1, go.z - coef_z * d2, go.w - coef_z * d3);
            }
            __syncwarp();
            if (lane == 0 && has_go) mbar_arrive(bar_empty + 8 * st);  // ring slot may be refilled
            if (it + 2 < niter) {
                named_sync(kGz, NGT);            // every gz warp is done reading this buffer
                stage(it + 2);                   // ... which tile it+2 re-uses
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < KD; i += 256) {
        const float v = acc[i];
        if (v != 0.0f) atomicAdd(&p.gE[i], coef_e * v);
    }
    peer_tail(p.peer, p.gE);
}

template <int D, int K, int HWT, int NST>
int launch_c1_tma(const BwdParams& p, cudaStream_t s) {
    constexpr int TM = 64;
    constexpr size_t smem = sizeof(float) * ((size_t)NST * D * TM + 2 * (size_t)TM + 2 * (size_t)D * (TM + 1) + 2 * (size_t)K * D) + (2 * NST + 2) * 8;
    static_assert(smem <= 227 * 1024, "shared memory");
    if (p.N % HWT != 0) return CTVQ_E_UNSUPPORTED;
    const long long nt = p.N / TM;  // one 64-position segment of an image per tile
    if (nt > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
    int grid = sm_count();
    if (grid > nt) grid = (int)nt;
    CUtensorMap gomap;
    memset(&gomap, 0, sizeof(gomap));
    if (HWT != TM && p.g_out != nullptr) {
        if (p.B > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
        const int rc = tc::make_plain_map(gomap, p.g_out, CTVQ_F32, HWT, D, p.B, TM, D);
        if (rc != CTVQ_OK) return rc;
    }
    auto kern = vq_bwd_c1_tma_kernel<D, K, HWT, NST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, 256, smem, s>>>(p, (int)nt, gomap);
    return (int)cudaGetLastError();
}
}  // namespace

int launch_backward_fast(const BwdParams& p, cudaStream_t s) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(p.gz) & 15) == 0) &&
                         (p.g_out == nullptr || (reinterpret_cast<uintptr_t>(p.g_out) & 15) == 0);
    if (!aligned && p.dtype == CTVQ_F32) return CTVQ_E_UNSUPPORTED;
    if (p.dtype == CTVQ_BF16) {
        // configs/mcq_vae.yaml shape with bf16 I/O: the TMA-ring kernel streams half the bytes
        const bool go16 = p.g_out == nullptr || (reinterpret_cast<uintptr_t>(p.g_out) & 127) == 0;
        if (p.d == 32 && p.K == 64 && p.cs == 1 && go16 && p.N % 64 == 0 && p.N >= (long long)sm_count() * 64 * 4 &&
            (reinterpret_cast<uintptr_t>(p.gz) & 7) == 0 && (reinterpret_cast<uintptr_t>(p.z) & 1) == 0) {
            if (p.C == 4 && p.Dtot == 128 && p.HW == 64) return launch_tma<32, 4, 64, 64, 128, 1, __nv_bfloat16>(p, s);
            if (p.C == 4 && p.Dtot == 128 && p.HW == 256) return launch_tma<32, 4, 64, 256, 128, 1, __nv_bfloat16>(p, s);
            if (p.C == 2 && p.Dtot == 64 && p.HW == 64) return launch_tma<32, 2, 64, 64, 64, 1, __nv_bfloat16>(p, s);
        }
        return CTVQ_E_UNSUPPORTED;  // the tiled kernel (ctvq_bwd.cu) carries bf16 for every other shape
    }
    static const bool no_tma = getenv("CTVQ_BWD_NO_TMA") != nullptr;  // A/B switch, read once per process
    // configs/mcq_vae.yaml: C=4, d=32, K=64, latents [B,128,8,8], overlapping slices
    if (p.d == 32 && p.C == 4 && p.K == 64 && p.HW == 64 && p.Dtot == 128 && p.cs == 1) {
        const bool go_ok = p.g_out == nullptr || (reinterpret_cast<uintptr_t>(p.g_out) & 127) == 0;
        if (go_ok && p.N >= (long long)sm_count() * 64 * 4 && !no_tma) return launch_tma<32, 4, 64, 64, 128, 1, float>(p, s);
        return launch<32, 4, 64, 64, 128, 1, 2>(p, s);
    }
    // neighbours of that shape (same kernel, one 64-position segment per tile): 128x128 images (H*W = 256) and two codebooks
    if (p.d == 32 && p.K == 64 && p.cs == 1 && (p.HW == 64 || p.HW == 256) && ((p.C == 4 && p.Dtot == 128) || (p.C == 2 && p.Dtot == 64)) &&
        (p.g_out == nullptr || (reinterpret_cast<uintptr_t>(p.g_out) & 127) == 0) && p.N >= (long long)sm_count() * 64 * 4 && !no_tma) {
        if (p.C == 4) return launch_tma<32, 4, 64, 256, 128, 1, float>(p, s);
        if (p.HW == 64) return launch_tma<32, 2, 64, 64, 64, 1, float>(p, s);
        return launch_tma<32, 2, 64, 256, 64, 1, float>(p, s);
    }
    // configs/ct_mcq_vae.yaml: C=1, d=128, K=64, latents [B,128,8,8]
    if (p.d == 128 && p.C == 1 && p.K == 64 && p.HW == 64 && p.Dtot == 128) {
        const bool ok = (reinterpret_cast<uintptr_t>(p.z) & 3) == 0 && (p.g_out == nullptr || (reinterpret_cast<uintptr_t>(p.g_out) & 127) == 0);
        if (ok && p.N >= (long long)sm_count() * 64 * 2 && !no_tma) return launch_c1_tma<128, 64, 64, 2>(p, s);
        return launch<128, 1, 64, 64, 128, 1, 1>(p, s);
    }
    // the same model on 128x128 images (H*W = 256): 64-position segments, g_out by one tensor-map box per tile
    if (p.d == 128 && p.C == 1 && p.K == 64 && p.HW == 256 && p.Dtot == 128 && (reinterpret_cast<uintptr_t>(p.z) & 3) == 0 &&
        (p.g_out == nullptr || (reinterpret_cast<uintptr_t>(p.g_out) & 127) == 0) && p.N >= (long long)sm_count() * 64 * 2 && !no_tma)
        return launch_c1_tma<128, 64, 256, 2>(p, s);
    return CTVQ_E_UNSUPPORTED;
}

}  // namespace ctvq
