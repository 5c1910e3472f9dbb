// Shape-specialised tcgen05 forward kernel for the multi-codebook quantiser (configs/mcq_vae.yaml shape).
// Same algorithm as ctvq_tc.cu (tf32 score GEMM in TMEM, rigorous candidate filter, exact fp32 re-scoring, fused
// gather / straight-through / loss) with every inner loop unrolled against compile-time D, NK, HW, C.
// Replaces models/vq_vae.py:30-55 / models/mcq_vae.py:26-64,100-127.
//
// One persistent CTA per SM, warp-specialised, NO CTA-wide barrier inside the tile loop:
//   producer warp (one lane)  TMA ring (cp.async.bulk.tensor.3d, NSTAGE tiles of 128 rows x 35 channels) and the
//                             tcgen05.mma group of each tile, issued up to two tiles ahead into a DOUBLE-BUFFERED
//                             accumulator (2 x 256 TMEM columns = all 512).  ALL C codebooks are ONE GEMM: the B operand
//                             is a [C*64 codes x 48] matrix holding codebook c at K-columns c*CS..c*CS+D-1 (zeros
//                             elsewhere), so the slab and the codebooks cross the shared-memory pipe once per tile;
//   16 epilogue warps         warpgroup g owns codebook g, warp (g, q) the 32 rows of TMEM lane quarter q.
// Hand-over is mbarrier-only: full/empty per ring slot, accumulator complete / drained per (buffer, codebook), so
// a warp that hits an expensive row (exact re-scoring) lags by itself instead of stalling the CTA.
//
// Per (row, code) work in the epilogue:
//     pass 1   s_k = z.e_k - |e_k|^2/2 comes straight out of TMEM (|e_k|^2 rides in the GEMM as one extra K-group:
//              A = constant ones, B = -|e_k|^2/2 split into three tf32 terms); 3-input max tree       (FMNMX3)
//     pass 2   survivor bitmask  s_k >= max - bound/2                         (FSETP + predicated LOP)
// Shared-memory bandwidth is the limiter (ncu: L1/shared pipe), so each thread reads its row from shared memory
// exactly once per tile into registers; |z|^2, the exact re-scoring and the gather/straight-through run from them.
#include "ctvq_tc_ptx.cuh"

namespace ctvq {
using namespace tc;
namespace {

__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU.SQRT, 2 ulp: inside the 1.0001 slack of the bound
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct FastParams {
    QuantParams q;
    int ntiles;
    unsigned long long* trace;  // debug: 8 globaltimer stamps per CTA (null in production), ctvq_debug_set_fast_trace()
};
unsigned long long* g_fast_trace = nullptr;

__device__ __forceinline__ void stamp(const FastParams& P, int slot) {
    if (P.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.trace[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

__device__ __forceinline__ void or_if_ge(unsigned& m, float a, float lim, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(m) : "f"(a), "f"(lim), "r"(bit));
}

// D: channels per codebook; NK: codes per codebook padded to 64; HWT: H*W; C: codebooks (C*NK <= 256);
// CS: channel stride between codebook slices (1 = the reference's overlapping slices); NSTAGE: TMA ring depth.
// ONE shared-memory slab of USEDP = round8((C-1)*CS + D) channels per row block serves every codebook: codebook c's
// UMMA descriptors simply start c*CS rows (128 B each) into it.
// NWG: epilogue warpgroups.  Work unit u = (tile u / C, codebook u % C); warpgroup g takes units g, g+NWG, ...
template <int D, int NK, int HWT, int C, int CS, int NSTAGE, int NWG>
__global__ void __launch_bounds__(128 * NWG + 32, 1) vq_fwd_tc_fast_kernel(const FastParams P, const __grid_constant__ Maps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const QuantParams& p = P.q;
    const int K = p.K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quarter = warp & 3, wg = warp >> 2;
    if (tid == 0) stamp(P, 0);
    constexpr int kFT = 128 * NWG + 32;                // NWG epilogue warpgroups + the producer warp
    constexpr int USEDP = ((C - 1) * CS + D + 7) / 8 * 8;
    constexpr int DJB = (D + 31) / 32;
    constexpr uint32_t kBlk = (uint32_t)USEDP * 128u;  // one 32-row block of the slab
    constexpr uint32_t kStage = 4u * kBlk;
    constexpr uint32_t kEcb = (uint32_t)DJB * NK * 128u;
    constexpr uint32_t kBblk = (uint32_t)C * NK * 128u;  // one K-block (32 K-columns) of the merged B operand
    static_assert(USEDP + 8 <= 64 && USEDP >= 32 && DJB == 1, "merged B operand: two K-blocks");
    static_assert(C * NK <= 256, "accumulator columns (two buffers fill the 512 TMEM columns)");
    static_assert(NK == 64 && C * 8 <= 32 && D % 8 == 0 && D % 4 == 0, "shape assumptions of this kernel");
    static_assert(NSTAGE >= 3, "the ring runs ahead of the double-buffered accumulator");
    uint8_t* a_s = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* e_s = a_s + (size_t)NSTAGE * kStage;
    // merged B operand of the score GEMM: [2 K-blocks][C*NK rows][32 floats], K-major SWIZZLE_128B.  Row 64c+k holds
    // E_c[k][j] at K-column c*CS+j, and -|e|^2/2 (three tf32 terms) at K-columns USEDP..USEDP+2 (the extra K-group,
    // whose A operand is the constant-ones block); everything else is zero.  e_s above is the plain per-codebook copy
    // the epilogue gathers from.
    uint8_t* b_s = e_s + (size_t)C * kEcb;
    uint8_t* ones_s = b_s + (size_t)2 * kBblk;  // [4 row blocks][8 channels][32 rows]: the A operand of the extra K-group
    float* ee_s = reinterpret_cast<float*>(ones_s + 4096);  // [C][NK]
    float* emax_s = ee_s + C * NK;                          // [C] (+pad)
    // mbarriers: full[NSTAGE] (TMA landed), empty[NSTAGE] (slot drained: every epilogue warp has its rows in registers
    // and the MMA group that read the slab is complete), mma[2] (accumulator buffer complete), tfree[2] (accumulator
    // buffer drained by all 4*C epilogue warps)
    uint64_t* bars = reinterpret_cast<uint64_t*>(emax_s + ((C + 3) & ~3));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
    const uint32_t a_base = smem_u32(a_s), b_base = smem_u32(b_s), ones_base = smem_u32(ones_s);
    const uint32_t bar_full0 = smem_u32(&bars[0]), bar_empty0 = smem_u32(&bars[NSTAGE]);
    const uint32_t bar_m = smem_u32(&bars[2 * NSTAGE]), bar_tfree = smem_u32(&bars[2 * NSTAGE + 2]);

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full0 + 8 * i, 1); mbar_init(bar_empty0 + 8 * i, 4 * C + 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_m + 8 * i, 1); mbar_init(bar_tfree + 8 * i, 4 * C); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    __syncthreads();  // barriers initialised before the first TMA may signal them
    if (tid == 0) stamp(P, 1);

    const int niter = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool producer = (tid == 128 * NWG);
    auto issue_tma = [&](int it) {  // producer only: TMA-load the tile of iteration `it` into its ring slot
        const int tile = blockIdx.x + it * gridDim.x;
        const int seg = tile / p.tiles_per_seg;
        const long long row0 = (long long)(tile - seg * p.tiles_per_seg) * kTM;
        const int st = it % NSTAGE;
        int nblk = 0;
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) nblk += (row0 + 32 * mb < p.N) ? 1 : 0;
        mbar_expect_tx(bar_full0 + 8 * st, (uint32_t)nblk * (uint32_t)((C - 1) * CS + D) * 128u);
        for (int mb = 0; mb < nblk; ++mb) {
            const long long nb = row0 + 32 * mb;
            const long long bb = nb / HWT;
            tma_load_3d(a_base + st * kStage + mb * kBlk, &maps.m[seg], bar_full0 + 8 * st, (int)(nb - bb * HWT), 0, (int)bb);
        }
    };
    if (producer)  // the first tiles stream in while the codebooks are staged
        for (int it = 0; it < NSTAGE && it < niter; ++it) issue_tma(it);

    // ---- codebooks (once per persistent CTA), one code per thread: the global loads fly while everybody zero-fills ----
    float4 v[D / 4];
    const int ck = tid % NK, cc = tid / NK;  // code, codebook of this thread (tid < C*NK)
    if (tid < C * NK) {
        if (ck < K) {
            const float4* row = reinterpret_cast<const float4*>(p.E[cc] + (size_t)ck * D);
#pragma unroll
            for (int m = 0; m < D / 4; ++m) v[m] = __ldg(row + m);  // all loads in flight at once
        } else {
#pragma unroll
            for (int m = 0; m < D / 4; ++m) v[m] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
    // zero the merged B operand, the slab channels no TMA box ever writes (they meet zero B columns, but 0 * garbage
    // could be NaN) and build the ones block
    for (int i = tid; i < (int)(2 * kBblk / 16); i += kFT) reinterpret_cast<float4*>(b_s)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    constexpr int USED = (C - 1) * CS + D;
    for (int i = tid; i < NSTAGE * 4 * (USEDP - USED) * 32; i += kFT) {
        const int col = i & 31, r = (i >> 5) % (USEDP - USED), blk = (i >> 5) / (USEDP - USED);
        reinterpret_cast<float*>(a_s + (size_t)blk * kBlk + (size_t)(USED + r) * 128)[col] = 0.0f;
    }
    for (int i = tid; i < 1024; i += kFT)  // rows are constant, so the swizzle inside a 128-byte row is immaterial
        reinterpret_cast<float*>(ones_s)[i] = ((i >> 5) & 7) < 3 ? 1.0f : 0.0f;
    if (tid < C) reinterpret_cast<unsigned*>(emax_s)[tid] = 0u;
    __syncthreads();
    if (tid < C * NK) {
        // plain per-codebook copy (the epilogue gathers from it), merged B at the shifted K-columns, exact |e|^2
        float a = 0.0f;  // exact sequential chain (arithmetic contract)
        const int n = cc * NK + ck;
#pragma unroll
        for (int m = 0; m < D / 4; ++m) {
            *reinterpret_cast<float4*>(e_s + (size_t)cc * kEcb + (m >> 3) * NK * 128 + ck * 128 + (((m & 7) ^ (ck & 7)) << 4)) = v[m];
            const float x[4] = {v[m].x, v[m].y, v[m].z, v[m].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = cc * CS + 4 * m + u;
                *reinterpret_cast<float*>(b_s + (size_t)(kk >> 5) * kBblk + e_off(n, kk & 31, C * NK)) = x[u];
                a = fmaf(x[u], x[u], a);
            }
        }
        if (ck >= K) a = CUDART_INF_F;
        ee_s[tid] = a;
        if (ck < K) atomicMax(reinterpret_cast<unsigned*>(emax_s) + cc, __float_as_uint(a));  // a >= 0: bit order = value order
        // -|e_k|^2/2 as three tf32-exact terms (30 mantissa bits) in the extra K-group; padded / overflowed codes get a
        // hugely negative score so they never survive the filter
        float t0 = -1.0e30f, t1 = 0.0f, t2 = 0.0f;
        if (a < CUDART_INF_F) {
            const float h = (-0.5f * kTruncC) * a;  // centred truncation error (ctvq_common.cuh)
            t0 = __uint_as_float(__float_as_uint(h) & 0xFFFFE000u);
            const float r1 = h - t0;
            t1 = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
            t2 = __uint_as_float(__float_as_uint(r1 - t1) & 0xFFFFE000u);
        }
        uint8_t* xb = b_s + (size_t)(USEDP >> 5) * kBblk;
        *reinterpret_cast<float*>(xb + e_off(tid, (USEDP & 31) + 0, C * NK)) = t0;
        *reinterpret_cast<float*>(xb + e_off(tid, (USEDP & 31) + 1, C * NK)) = t1;
        *reinterpret_cast<float*>(xb + e_off(tid, (USEDP & 31) + 2, C * NK)) = t2;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid < C) emax_s[tid] = sqrtf(__uint_as_float(reinterpret_cast<unsigned*>(emax_s)[tid])) * 1.0001f;
    const uint32_t tmem_base = *tmem_slot;
    __syncthreads();
    if (tid == 0) stamp(P, 2);

    float lsum[C];
#pragma unroll
    for (int i = 0; i < C; ++i) lsum[i] = 0.0f;
    unsigned nnear = 0u;  // near-tie rows seen by this thread (include/ctvq.h, CTVQ_NEAR_TIE_REL)
    if (warp == 4 * NWG) {
        // =============================== producer: TMA ring + MMA groups ===========================================
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32(C * NK);
            int tma_next = NSTAGE < niter ? NSTAGE : niter;
            for (int it = 0; it < niter; ++it) {
                const int st = it % NSTAGE, buf = it & 1;
                const uint32_t stage_u32 = a_base + st * kStage;
                mbar_wait_sleep(bar_full0 + 8 * st, (uint32_t)(it / NSTAGE) & 1u);
                // accumulator buffer `buf` was last read for tile it-2
                if (it >= 2) mbar_wait_sleep(bar_tfree + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u);
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * 256;
#pragma unroll
                for (int s = 0; s < USEDP / 8; ++s) {  // one GEMM for all codebooks: [128 rows x USEDP] x [USEDP x C*NK]
                    const uint64_t ad = smem_desc(stage_u32 + (uint32_t)(8 * s) * 128u, kBlk, 512u, 1u);
                    const uint64_t bd = smem_desc(b_base + (s >> 2) * kBblk + (s & 3) * 32u, 16u, 1024u, 2u);
                    umma_tf32(dcol, ad, bd, idesc, s > 0 ? 1u : 0u);
                }
                // + 1 * (-|e_k|^2 / 2): the accumulator now holds the whole score z.e_k - |e_k|^2/2
                umma_tf32(dcol, smem_desc(ones_base, 1024u, 512u, 1u),
                          smem_desc(b_base + ((USEDP / 8) >> 2) * kBblk + ((USEDP / 8) & 3) * 32u, 16u, 1024u, 2u), idesc, 1u);
                umma_commit(bar_m + 8 * buf);
                umma_commit(bar_empty0 + 8 * st);  // the same MMAs were the last readers of the slab
                // top the ring up: slot of tile tma_next was last used by tile tma_next - NSTAGE
                while (tma_next < niter && tma_next <= it + NSTAGE - 1) {
                    const int prev = tma_next - NSTAGE;
                    mbar_wait_sleep(bar_empty0 + 8 * (prev % NSTAGE), (uint32_t)(prev / NSTAGE) & 1u);
                    issue_tma(tma_next);
                    ++tma_next;
                }
            }
        }
    } else {
        // =============================== epilogue warps: units (tile, codebook), rows of lane quarter ==================
        const int nunits = niter * C;
        // one warpgroup per codebook (NWG == C): the codebook of a warpgroup never changes, so everything that depends on it
        // only is computed once (the compiler cannot see that u % C is loop-invariant)
        constexpr bool kFixedC = (NWG == C);
        int c = kFixedC ? wg : 0;
        // lane-dependent part of the swizzled z address, indexed by (channel & 3) relative to this codebook's first channel
        uint32_t zsw[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) zsw[x] = ((((lane >> 3) ^ (x + c * CS)) & 3) << 5) + ((lane & 7) << 2);
        const float* ee = ee_s + c * NK;
        const uint8_t* ecb = e_s + (size_t)c * kEcb;
        float emax = emax_s[c];
        const bool one_seg = (p.n_seg == 1);
        for (int u = wg; u < nunits; u += NWG) {
            const int it = kFixedC ? (u - wg) / NWG : u / C;
            if (!kFixedC) {
                c = u - it * C;
#pragma unroll
                for (int x = 0; x < 4; ++x) zsw[x] = ((((lane >> 3) ^ (x + c * CS)) & 3) << 5) + ((lane & 7) << 2);
                ee = ee_s + c * NK;
                ecb = e_s + (size_t)c * kEcb;
                emax = emax_s[c];
            }
            const int tile = blockIdx.x + it * gridDim.x;
            const int seg = one_seg ? 0 : tile / p.tiles_per_seg;  // (a runtime division per tile otherwise)
            const long long row0 = (long long)(tile - seg * p.tiles_per_seg) * kTM;
            const long long n = row0 + quarter * 32 + lane;
            const bool valid = n < p.N;  // warp-uniform (N is a multiple of 32)
            const long long b = n / HWT;
            const int hw = (int)(n - b * HWT);
            const int st = it % NSTAGE, buf = it & 1;
            mbar_wait_fast(bar_full0 + 8 * st, (uint32_t)(it / NSTAGE) & 1u);
            if (tid == 0 && u == 0) stamp(P, 3);    // first slab landed
            if (tid == 0 && u == NWG) stamp(P, 4);  // first unit of this warp finished
            // this thread's row, every channel of its codebook, read from shared memory ONCE into registers
            const uint8_t* zblk = a_s + st * kStage + quarter * kBlk + c * CS * 128;
            float zr[D];
            if (valid) {
#pragma unroll
                for (int j = 0; j < D; ++j) zr[j] = *reinterpret_cast<const float*>(zblk + j * 128 + zsw[j & 3]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty0 + 8 * st);  // this warp no longer needs the slab
            mbar_wait_fast(bar_m + 8 * buf, (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256 + c * NK;
            float mx = 0.0f, zzc = 0.0f;
            unsigned mask0 = 0u, mask1 = 0u;
            if (valid) {
                uint32_t a[32];
                // ---- pass 1: approximate scores s_k = z.e_k - |e_k|^2/2 (= -(dist_k - |z|^2)/2) and their maximum -----------
                // (two 32-column halves, re-read from TMEM in pass 2: 32 live registers instead of 64); the exact sequential
                // |z|^2 chain (arithmetic contract) runs underneath the first TMEM load
                float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
                tmem_ld32_issue(trow, a);
#pragma unroll
                for (int j = 0; j < D; ++j) zzc = fmaf(zr[j], zr[j], zzc);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 1) tmem_ld32_issue(trow + 32, a);
                    tmem_ld32_wait(a);
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        m0 = fmaxf(m0, __uint_as_float(a[i])); m1 = fmaxf(m1, __uint_as_float(a[i + 1]));
                        m2 = fmaxf(m2, __uint_as_float(a[i + 2])); m3 = fmaxf(m3, __uint_as_float(a[i + 3]));
                    }
                }
                mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                // rigorous bound on |tf32 distance - exact-chain distance|: both operands lose at most
                // 2^-10 relative (truncation to 10 mantissa bits) -> |dot error| <= (2^-9 + slack) |z||e|; scores are
                // distances / -2, so the window is half the distance bound
                const float thr = 2.0f * (2.0f * kTf32Eps * sqrt_approx(zzc) * 1.0001f * emax + kWinAbs * (zzc + emax * emax));
                const float lim = mx - 0.5f * thr;
                // ---- pass 2: survivors as a bitmask (four independent accumulators per half); the upper half is still in
                // registers from pass 1, only the lower half is re-read from TMEM
#pragma unroll
                for (int h = 1; h >= 0; --h) {
                    if (h == 0) {
                        tmem_ld32_issue(trow, a);
                        tmem_ld32_wait(a);
                    }
                    unsigned mk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < 32; ++i) or_if_ge(mk[i & 3], __uint_as_float(a[i]), lim, 1u << i);
                    const unsigned m = (mk[0] | mk[1]) | (mk[2] | mk[3]);
                    if (h == 0) mask0 = m; else mask1 = m;
                }
            }
            // the accumulator is drained: hand it back to the producer for tile it+2
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfree + 8 * buf);
            if (valid) {
                const int cnt = __popc(mask0) + __popc(mask1);
                // ---- decide ---------------------------------------------------------------------------------------------------
                int bi = 0;
                const bool finite = (zzc < CUDART_INF_F) && (mx > -CUDART_INF_F) && (mx < CUDART_INF_F) && cnt >= 1;
                if (finite && cnt == 1) {
                    bi = mask0 ? __ffs(mask0) - 1 : 32 + __ffs(mask1) - 1;
                } else {
                    float bv = CUDART_INF_F, bv2 = CUDART_INF_F;  // exact best / second-best distance
                    bi = 0x7fffffff;
                    if (!finite) {
                        // non-finite row: exact scan of every code with torch.argmin's NaN rule
                        for (int k = 0; k < K; ++k) {
                            const uint8_t* erow = ecb + k * 128;
                            float dot = 0.0f;
#pragma unroll
                            for (int j = 0; j < D; ++j)
                                dot = fmaf(zr[j], *reinterpret_cast<const float*>(erow + (j >> 5) * NK * 128 +
                                                                                   (((((j & 31) >> 2) ^ (k & 7)) & 7) << 4) + ((j & 3) << 2)), dot);
                            const float dist = dist_f32(zzc, ee[k], dot);
                            if (k == 0 || (!(dist >= bv) && (bv == bv))) { bv = dist; bi = k; }  // k == 0 seeds the scan: an all-+inf row answers 0 like torch.argmin
                        }
                    } else {
                        // exact re-scoring of the survivors, ascending k over the whole 64-bit mask (one loop: the trip count is
                        // the warp's largest survivor count)
                        unsigned long long mk = ((unsigned long long)mask1 << 32) | mask0;
                        while (mk) {  // two candidates per trip: both code rows in flight, the two FMA chains interleaved
                            const int ka = __ffsll((long long)mk) - 1;
                            mk &= mk - 1;
                            const int kb = mk ? __ffsll((long long)mk) - 1 : ka;
                            mk &= mk - 1;
                            const uint8_t* ra = ecb + ka * 128;
                            const uint8_t* rb = ecb + kb * 128;
                            const uint32_t xa = (uint32_t)(ka & 7) << 4, xb = (uint32_t)(kb & 7) << 4;
                            float da = 0.0f, db = 0.0f;
#pragma unroll
                            for (int j = 0; j < D; j += 4) {
                                const float4 a4 = *reinterpret_cast<const float4*>(ra + (j >> 5) * NK * 128 + ((((j & 31) >> 2) << 4) ^ xa));
                                const float4 b4 = *reinterpret_cast<const float4*>(rb + (j >> 5) * NK * 128 + ((((j & 31) >> 2) << 4) ^ xb));
                                da = fmaf(zr[j], a4.x, da); db = fmaf(zr[j], b4.x, db);
                                da = fmaf(zr[j + 1], a4.y, da); db = fmaf(zr[j + 1], b4.y, db);
                                da = fmaf(zr[j + 2], a4.z, da); db = fmaf(zr[j + 2], b4.z, db);
                                da = fmaf(zr[j + 3], a4.w, da); db = fmaf(zr[j + 3], b4.w, db);
                            }
                            const float dista = dist_f32(zzc, ee[ka], da), distb = dist_f32(zzc, ee[kb], db);
                            if (dista < bv) { bv2 = bv; bv = dista; bi = ka; }  // ascending k: strict '<' keeps the first minimum
                            else bv2 = fminf(bv2, dista);
                            if (kb != ka) {  // (kb == ka when the mask ran out)
                                if (distb < bv) { bv2 = bv; bv = distb; bi = kb; }
                                else bv2 = fminf(bv2, distb);
                            }
                        }
                        nnear += near_tie(bv, bv2) ? 1u : 0u;  // the window holds the exact runner-up of every near-tie row (kWinAbs)
                    }
                }
                p.idx[seg][((size_t)b * C + c) * HWT + hw] = (long long)bi;
                // ---- fused gather + straight-through + loss ------------------------------------------------------------------
                if (p.fused) {
                    float* out = p.q + ((size_t)b * C * D + (size_t)c * D) * HWT + hw;  // a warp's 32 rows are contiguous: 128-byte stores
                    const uint8_t* erow = ecb + bi * 128;
                    const uint32_t kx = (uint32_t)(bi & 7) << 4;
                    float ls0 = 0.0f, ls1 = 0.0f;
#pragma unroll
                    for (int j = 0; j < D; j += 4) {
                        const float4 e4 = *reinterpret_cast<const float4*>(erow + (j >> 5) * NK * 128 + ((((j & 31) >> 2) << 4) ^ kx));
                        const float d0 = __fsub_rn(e4.x, zr[j]), d1 = __fsub_rn(e4.y, zr[j + 1]);
                        const float d2 = __fsub_rn(e4.z, zr[j + 2]), d3 = __fsub_rn(e4.w, zr[j + 3]);
                        out[(size_t)j * HWT] = __fadd_rn(zr[j], d0);  // z + (q - z), models/vq_vae.py:53
                        out[(size_t)(j + 1) * HWT] = __fadd_rn(zr[j + 1], d1);
                        out[(size_t)(j + 2) * HWT] = __fadd_rn(zr[j + 2], d2);
                        out[(size_t)(j + 3) * HWT] = __fadd_rn(zr[j + 3], d3);
                        ls0 = fmaf(d0, d0, ls0); ls1 = fmaf(d1, d1, ls1);
                        ls0 = fmaf(d2, d2, ls0); ls1 = fmaf(d3, d3, ls1);
                    }
                    const float ls = ls0 + ls1;
#pragma unroll
                    for (int i = 0; i < C; ++i) lsum[i] += (i == c) ? ls : 0.0f;  // static register indexing
                }
            }
        }
    }
    if (tid == 0) stamp(P, 5);  // this warp's last unit done
    if (p.neartie && warp < 4 * NWG) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, nnear);
        if (lane == 0 && tot) atomicAdd(p.neartie, (unsigned long long)tot);
    }
    // ---- loss: warp sums -> fp64 atomics -> last CTA finalises -----------------------------------------------------
    if (p.fused) {
        if (warp < 4 * NWG) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                double v = (double)lsum[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && v != 0.0) atomicAdd(&p.loss_acc[c], v);
            }
        }
        __shared__ unsigned s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1u);
        __syncthreads();
        if (s_last && tid == 0) {
            __threadfence();
            float total = 0.0f;
            const double denom = (double)p.N * (double)D;
            for (int c = 0; c < C; ++c) {
                const float m = (float)(__ldcg(&p.loss_acc[c]) / denom);
                const float l = __fadd_rn(__fmul_rn(m, p.beta), m);
                p.loss_out[c] = l;
                total = __fadd_rn(total, l);
                p.loss_acc[c] = 0.0;
            }
            p.loss_out[C] = total;
            *p.ticket = 0u;
            __threadfence();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) stamp(P, 6);
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int D, int NK, int HWT, int C, int CS, int NSTAGE, int NWG>
int launch_fast(const QuantParams& p0, cudaStream_t s) {
    FastParams P;
    P.q = p0;
    P.q.tiles_per_seg = (int)((p0.N + kTM - 1) / kTM);
    P.ntiles = P.q.tiles_per_seg * p0.n_seg;
    P.trace = g_fast_trace;
    constexpr int USEDP = ((C - 1) * CS + D + 7) / 8 * 8;
    constexpr int DJB = (D + 31) / 32;
    Maps maps;
    if (make_maps(p0, maps, (C - 1) * CS + D) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;  // rows USED..USEDP-1 of the slab stay unwritten and unread
    constexpr size_t smem = (size_t)NSTAGE * 4 * USEDP * 128 + (size_t)C * DJB * NK * 128 + (size_t)2 * C * NK * 128 + 4096 +
                            sizeof(float) * ((size_t)C * NK + ((C + 3) & ~3)) + (2 * NSTAGE + 4) * 8 + 16 + 1024;
    static_assert(smem <= 227 * 1024, "one CTA per SM");
    auto kern = vq_fwd_tc_fast_kernel<D, NK, HWT, C, CS, NSTAGE, NWG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int grid = sm_count();
    if (grid > P.ntiles) grid = P.ntiles;
    kern<<<grid, 128 * NWG + 32, smem, s>>>(P, maps);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" void ctvq_debug_set_fast_trace(unsigned long long* buf) { g_fast_trace = buf; }  // [SMs*8] device words or null

// Shapes with a specialised kernel; anything else falls through to the generic tcgen05 kernel / SIMT.
int launch_forward_tc_fast(const QuantParams& p, cudaStream_t s) {
    if (p.HW % 32 != 0 || p.K > 64) return CTVQ_E_UNSUPPORTED;
    for (int sg = 0; sg < p.n_seg; ++sg)
        if (reinterpret_cast<uintptr_t>(p.z[sg]) & 15) return CTVQ_E_UNSUPPORTED;
    for (int c = 0; c < p.C; ++c)
        if (reinterpret_cast<uintptr_t>(p.E[c]) & 15) return CTVQ_E_UNSUPPORTED;  // 128-bit codebook loads
    if (!encode_fn()) return CTVQ_E_UNSUPPORTED;
    // configs/mcq_vae.yaml: C=4 codebooks x d=32 on overlapping slices of [B,128,8,8]
    if (p.d == 32 && p.HW == 64 && p.C == 4 && p.cs == 1) {
        // 4 epilogue warpgroups: 0.160 ms at 1 M rows when measured (0.152 ms today); 5 / 6 warpgroups (80 / 72 registers) measured 0.176 / 0.180 ms
        // -- the L1/shared data pipe, not latency, is the limiter, so more warps only add contention
        return launch_fast<32, 64, 64, 4, 1, 5, 4>(p, s);
    }
    // neighbours of that shape: the same model on 128x128 images (H*W = 256), and two codebooks instead of four
    if (p.d == 32 && p.cs == 1 && p.C == 4 && p.HW == 256) return launch_fast<32, 64, 256, 4, 1, 5, 4>(p, s);
    if (p.d == 32 && p.cs == 1 && p.C == 2 && p.HW == 64) return launch_fast<32, 64, 64, 2, 1, 5, 4>(p, s);
    if (p.d == 32 && p.cs == 1 && p.C == 2 && p.HW == 256) return launch_fast<32, 64, 256, 2, 1, 5, 4>(p, s);
    return CTVQ_E_UNSUPPORTED;
}

}  // namespace ctvq
