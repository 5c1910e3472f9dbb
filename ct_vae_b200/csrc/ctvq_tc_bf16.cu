// bf16-latent kernels for the multi-codebook quantiser shape of configs/mcq_vae.yaml (C=4, d=32, K=64, [B,128,8,8]):
// dtype = CTVQ_BF16 of include/ctvq.h.  The reference defines no bf16 mode (`.bfloat16()` raises at models/vq_vae.py:43);
// ours is the fp32 arithmetic contract of DESIGN.md applied to bf16-ROUNDED latents and codebooks, outputs rounded to
// bf16 when stored.  Same maths as models/vq_vae.py:30-55 / models/mcq_vae.py:26-64,100-127.
//
// Forward = vq_fwd_tc_fast_kernel (ctvq_tc_fast.cu) re-cut for 16-bit operands:
//   * the NCHW slab arrives by TMA as bf16 (HALF the HBM and shared-memory bytes): one [35 ch x 64 rows] box per 64-row
//     block lands as a 128-byte-row MN-major SWIZZLE_128B operand (64 rows = one swizzle row);
//   * tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32 in TMEM), K = 16 per instruction: THREE MMAs per tile instead of six.
//     bf16 products are exact in fp32, so the only error of the tensor-core score is its accumulation: the candidate
//     window shrinks from ~4e-3 |z||e| (tf32 truncation) to ~1e-5 (|z|^2 + |e|^2) and exact re-scoring all but
//     disappears (only genuine near-ties survive the filter);
//   * |e|^2 still rides in the GEMM: the merged B operand carries -|e_k|^2/2 as FOUR bf16 terms (32 mantissa bits) at
//     K-columns 40..43, against constant-one rows 40..43 of every slab block.
// Backward = vq_bwd_tma_kernel<..., __nv_bfloat16> (ctvq_bwd_fast.cu) at this shape, the tiled kernel (ctvq_bwd.cu) elsewhere.
#include "ctvq_tc_ptx.cuh"

namespace ctvq {
using namespace tc;
namespace {

__device__ __forceinline__ float sqrt_approx_h(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void or_if_ge_h(unsigned& m, float a, float lim, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(m) : "f"(a), "f"(lim), "r"(bit));
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16, bf16 x bf16 -> fp32, A MN-major (rows contiguous), B K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
}
// bf16 element (row r of a 64-row block, channel j) inside an A block [KP][128 B]: TMA SWIZZLE_128B = cute Swizzle<3,4,3>
__device__ __forceinline__ uint32_t a16_off(int r, int j) {
    return (uint32_t)(j * 128 + ((((r >> 3) ^ j) & 7) << 4) + ((r & 7) << 1));
}
// bf16 element (row n, K-column kk < 64) of the merged B operand [rows][128 B], K-major SWIZZLE_128B
__device__ __forceinline__ uint32_t b16_off(int n, int kk) {
    return (uint32_t)(n * 128 + ((((kk >> 3) ^ n) & 7) << 4) + ((kk & 7) << 1));
}
__device__ __forceinline__ unsigned short bf16_bits_rn(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }

// Absolute term of the candidate window for exact (bf16 x bf16) products: only fp32 accumulation inside the tensor core
// and the fp32 rounding of the distance formula separate the tensor-core score from the exact-chain distance; 8 x 2^-20
// of (|z|^2 + max|e|^2) bounds 48 products summed with at least 22 bits kept, plus the near-tie widening of kWinAbs.
constexpr float kWinAbs16 = 8.0f * 9.5367431640625e-7f + 1.0e-6f;

struct Bf16Params {
    QuantParams q;
    int ntiles;
};

// D: channels per codebook; NK: codes per codebook padded to 64; HWT: H*W (a multiple of 64); C: codebooks (C*NK <= 256);
// CS: channel stride between codebook slices; NSTAGE: TMA ring depth; NWG: epilogue warpgroups.
template <int D, int NK, int HWT, int C, int CS, int NSTAGE, int NWG>
__global__ void __launch_bounds__(128 * NWG + 32, 1) vq_fwd_tc_bf16_kernel(const Bf16Params P, const __grid_constant__ Maps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const QuantParams& p = P.q;
    const int K = p.K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quarter = warp & 3, wg = warp >> 2;
    constexpr int kFT = 128 * NWG + 32;
    constexpr int USED = (C - 1) * CS + D;            // channels the slices touch (TMA box)
    constexpr int XR = (USED + 7) / 8 * 8;            // first constant-one row: the -|e|^2/2 terms' K-columns
    constexpr int KP = (XR + 4 + 15) / 16 * 16;       // K extent of the GEMM (multiple of the kind::f16 K = 16)
    constexpr uint32_t kBlk = (uint32_t)KP * 128u;    // one 64-row block of the slab: [KP][128 B]
    constexpr uint32_t kStage = 2u * kBlk;            // 128 rows
    constexpr uint32_t kEcb = (uint32_t)NK * 128u;    // plain fp32 copy of one codebook (rows of 32 floats)
    constexpr uint32_t kB = (uint32_t)C * NK * 128u;  // merged B operand: [C*NK rows][64 bf16]
    static_assert(KP <= 64 && D == 32, "one 128-byte K-block of bf16; rows of the fp32 copy are 32 floats");
    static_assert(C * NK <= 256 && NK == 64 && HWT % 64 == 0, "shape assumptions of this kernel");
    static_assert(NSTAGE >= 3, "the ring runs ahead of the double-buffered accumulator");
    uint8_t* a_s = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* e_s = a_s + (size_t)NSTAGE * kStage;     // [C][NK][32 floats], 16-byte chunks XOR-swizzled by (k & 7): gather / re-scoring
    uint8_t* b_s = e_s + (size_t)C * kEcb;
    float* ee_s = reinterpret_cast<float*>(b_s + kB);  // [C][NK] exact |e|^2 of the ROUNDED codebook
    float* emax_s = ee_s + C * NK;                      // [C] (+pad)
    uint64_t* bars = reinterpret_cast<uint64_t*>(emax_s + ((C + 3) & ~3));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
    unsigned* s_last = tmem_slot + 1;
    const uint32_t a_base = smem_u32(a_s), b_base = smem_u32(b_s);
    const uint32_t bar_full0 = smem_u32(&bars[0]), bar_empty0 = smem_u32(&bars[NSTAGE]);
    const uint32_t bar_m = smem_u32(&bars[2 * NSTAGE]), bar_tfree = smem_u32(&bars[2 * NSTAGE + 2]);

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full0 + 8 * i, 1); mbar_init(bar_empty0 + 8 * i, 4 * C + 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_m + 8 * i, 1); mbar_init(bar_tfree + 8 * i, 4 * C); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    __syncthreads();

    const int niter = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool producer = (tid == 128 * NWG);
    auto issue_tma = [&](int it) {  // producer only: TMA-load the tile of iteration `it` (two 64-row blocks) into its ring slot
        const int tile = blockIdx.x + it * gridDim.x;
        const int seg = tile / p.tiles_per_seg;
        const long long row0 = (long long)(tile - seg * p.tiles_per_seg) * kTM;
        const int st = it % NSTAGE;
        int nblk = 0;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) nblk += (row0 + 64 * mb < p.N) ? 1 : 0;
        mbar_expect_tx(bar_full0 + 8 * st, (uint32_t)nblk * (uint32_t)USED * 128u);
        for (int mb = 0; mb < nblk; ++mb) {
            const long long nb = row0 + 64 * mb;
            const long long bb = nb / HWT;
            tma_load_3d(a_base + st * kStage + mb * kBlk, &maps.m[seg], bar_full0 + 8 * st, (int)(nb - bb * HWT), 0, (int)bb);
        }
    };
    if (producer)
        for (int it = 0; it < NSTAGE && it < niter; ++it) issue_tma(it);

    // ---- codebooks (once per persistent CTA), one code per thread, ROUNDED to bf16 as they are read ---------------------
    float v[D];
    const int ck = tid % NK, cc = tid / NK;  // code, codebook of this thread (tid < C*NK)
    if (tid < C * NK) {
        if (ck < K) {
            const float4* row = reinterpret_cast<const float4*>(p.E[cc] + (size_t)ck * D);
#pragma unroll
            for (int m = 0; m < D / 4; ++m) {
                const float4 t = __ldg(row + m);
                v[4 * m] = IO<__nv_bfloat16>::cb(t.x); v[4 * m + 1] = IO<__nv_bfloat16>::cb(t.y);
                v[4 * m + 2] = IO<__nv_bfloat16>::cb(t.z); v[4 * m + 3] = IO<__nv_bfloat16>::cb(t.w);
            }
        } else {
#pragma unroll
            for (int m = 0; m < D; ++m) v[m] = 0.0f;
        }
    }
    for (int i = tid; i < (int)(kB / 16); i += kFT) reinterpret_cast<float4*>(b_s)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    // slab rows no TMA box ever writes: zeros, except the four constant-one rows XR..XR+3 (rows are constant: swizzle immaterial)
    for (int i = tid; i < NSTAGE * 2 * (KP - USED) * 64; i += kFT) {
        const int col = i & 63, r = (i >> 6) % (KP - USED), blk = (i >> 6) / (KP - USED);
        const int row = USED + r;
        reinterpret_cast<unsigned short*>(a_s + (size_t)blk * kBlk + (size_t)row * 128)[col] = (row >= XR && row < XR + 4) ? 0x3F80u : 0u;
    }
    if (tid < C) reinterpret_cast<unsigned*>(emax_s)[tid] = 0u;
    __syncthreads();
    if (tid < C * NK) {
        float a = 0.0f;  // exact sequential chain (arithmetic contract) over the ROUNDED values
        const int n = cc * NK + ck;
#pragma unroll
        for (int m = 0; m < D / 4; ++m)
            *reinterpret_cast<float4*>(e_s + (size_t)cc * kEcb + ck * 128 + (((m & 7) ^ (ck & 7)) << 4)) =
                make_float4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
#pragma unroll
        for (int j = 0; j < D; ++j) {
            *reinterpret_cast<unsigned short*>(b_s + b16_off(n, cc * CS + j)) = bf16_bits_rn(v[j]);  // exact: v is bf16-valued
            a = fmaf(v[j], v[j], a);
        }
        if (ck >= K) a = CUDART_INF_F;
        ee_s[tid] = a;
        if (ck < K) atomicMax(reinterpret_cast<unsigned*>(emax_s) + cc, __float_as_uint(a));
        // -|e_k|^2/2 as four bf16-exact terms (truncation: each remainder is exact in fp32); padded / overflowed codes get a
        // hugely negative score so they never survive the filter
        float t[4] = {-1.0e30f, 0.0f, 0.0f, 0.0f};
        if (a < CUDART_INF_F) {
            float r = -0.5f * a;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                t[i] = __uint_as_float(__float_as_uint(r) & 0xFFFF0000u);
                r -= t[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<unsigned short*>(b_s + b16_off(n, XR + i)) = (unsigned short)(__float_as_uint(t[i]) >> 16);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid < C) emax_s[tid] = sqrtf(__uint_as_float(reinterpret_cast<unsigned*>(emax_s)[tid])) * 1.0001f;
    const uint32_t tmem_base = *tmem_slot;
    __syncthreads();

    float lsum[C];
#pragma unroll
    for (int i = 0; i < C; ++i) lsum[i] = 0.0f;
    unsigned nnear = 0u;
    if (warp == 4 * NWG) {
        // =============================== producer: TMA ring + MMA groups ===========================================
        if (lane == 0) {
            const uint32_t idesc = instr_desc_bf16(C * NK);
            int tma_next = NSTAGE < niter ? NSTAGE : niter;
            for (int it = 0; it < niter; ++it) {
                const int st = it % NSTAGE, buf = it & 1;
                const uint32_t stage_u32 = a_base + st * kStage;
                mbar_wait_sleep(bar_full0 + 8 * st, (uint32_t)(it / NSTAGE) & 1u);
                if (it >= 2) mbar_wait_sleep(bar_tfree + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u);
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * 256;
#pragma unroll
                for (int s = 0; s < KP / 16; ++s) {  // [128 rows x KP] x [KP x C*NK], 16 K-columns per instruction
                    // A: MN-major SWIZZLE_128B, 64-row blocks kBlk apart (LBO), 8-channel groups 1 KB apart (SBO)
                    const uint64_t ad = smem_desc(stage_u32 + (uint32_t)s * 2048u, kBlk, 1024u, 2u);
                    // B: K-major SWIZZLE_128B, 8-row groups 1 KB apart (SBO); 16 K-columns = 32 bytes into the 128-byte row
                    const uint64_t bd = smem_desc(b_base + (uint32_t)s * 32u, 16u, 1024u, 2u);
                    umma_bf16(dcol, ad, bd, idesc, s > 0 ? 1u : 0u);
                }
                umma_commit(bar_m + 8 * buf);
                umma_commit(bar_empty0 + 8 * st);
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
        const int rblk = (quarter & 1) * 32 + lane;  // row inside its 64-row block
        constexpr bool kFixedC = (NWG == C);  // one warpgroup per codebook: c is loop-invariant (ctvq_tc_fast.cu)
        const bool one_seg = (p.n_seg == 1);
        for (int u = wg; u < nunits; u += NWG) {
            const int it = kFixedC ? (u - wg) / NWG : u / C;
            const int c = kFixedC ? wg : u - it * C;
            const float* ee = ee_s + c * NK;
            const uint8_t* ecb = e_s + (size_t)c * kEcb;
            const float emax = emax_s[c];
            const int tile = blockIdx.x + it * gridDim.x;
            const int seg = one_seg ? 0 : tile / p.tiles_per_seg;  // (a runtime division per tile otherwise)
            const long long row0 = (long long)(tile - seg * p.tiles_per_seg) * kTM;
            const long long n = row0 + quarter * 32 + lane;
            const bool valid = n < p.N;  // warp-uniform (N is a multiple of 32)
            const long long b = n / HWT;
            const int hw = (int)(n - b * HWT);
            const int st = it % NSTAGE, buf = it & 1;
            mbar_wait_fast(bar_full0 + 8 * st, (uint32_t)(it / NSTAGE) & 1u);
            // this thread's row, every channel of its codebook, read from shared memory ONCE into registers (bf16 -> fp32: exact)
            const uint8_t* zblk = a_s + st * kStage + (quarter >> 1) * kBlk;
            float zr[D];
            if (valid) {
#pragma unroll
                for (int j = 0; j < D; ++j)
                    zr[j] = __uint_as_float((unsigned)*reinterpret_cast<const unsigned short*>(zblk + a16_off(rblk, c * CS + j)) << 16);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty0 + 8 * st);
            mbar_wait_fast(bar_m + 8 * buf, (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256 + c * NK;
            float mx = 0.0f, zzc = 0.0f;
            unsigned mask0 = 0u, mask1 = 0u;
            if (valid) {
                uint32_t a[32];
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
                // products are exact: the window only covers fp32 accumulation (see kWinAbs16); scores are distances / -2
                const float thr = 2.0f * kWinAbs16 * (zzc + emax * emax);
                const float lim = mx - 0.5f * thr;
#pragma unroll
                for (int h = 1; h >= 0; --h) {
                    if (h == 0) {
                        tmem_ld32_issue(trow, a);
                        tmem_ld32_wait(a);
                    }
                    unsigned mk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < 32; ++i) or_if_ge_h(mk[i & 3], __uint_as_float(a[i]), lim, 1u << i);
                    const unsigned m = (mk[0] | mk[1]) | (mk[2] | mk[3]);
                    if (h == 0) mask0 = m; else mask1 = m;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfree + 8 * buf);
            if (valid) {
                const int cnt = __popc(mask0) + __popc(mask1);
                int bi = 0;
                const bool finite = (zzc < CUDART_INF_F) && (mx > -CUDART_INF_F) && (mx < CUDART_INF_F) && cnt >= 1;
                if (finite && cnt == 1) {
                    bi = mask0 ? __ffs(mask0) - 1 : 32 + __ffs(mask1) - 1;
                } else {
                    float bv = CUDART_INF_F, bv2 = CUDART_INF_F;
                    bi = 0x7fffffff;
                    if (!finite) {
                        for (int k = 0; k < K; ++k) {  // non-finite row: exact scan with torch.argmin's NaN rule
                            const uint8_t* erow = ecb + k * 128;
                            float dot = 0.0f;
#pragma unroll
                            for (int j = 0; j < D; ++j)
                                dot = fmaf(zr[j], *reinterpret_cast<const float*>(erow + ((((j >> 2) ^ (k & 7)) & 7) << 4) + ((j & 3) << 2)), dot);
                            const float dist = dist_f32(zzc, ee[k], dot);
                            if (k == 0 || (!(dist >= bv) && (bv == bv))) { bv = dist; bi = k; }
                        }
                    } else {
                        unsigned long long mk = ((unsigned long long)mask1 << 32) | mask0;
                        while (mk) {  // exact re-scoring of the survivors, ascending k (rare here: genuine near-ties only)
                            const int ka = __ffsll((long long)mk) - 1;
                            mk &= mk - 1;
                            const uint8_t* ra = ecb + ka * 128;
                            const uint32_t xa = (uint32_t)(ka & 7) << 4;
                            float da = 0.0f;
#pragma unroll
                            for (int j = 0; j < D; j += 4) {
                                const float4 a4 = *reinterpret_cast<const float4*>(ra + (((j >> 2) << 4) ^ xa));
                                da = fmaf(zr[j], a4.x, da); da = fmaf(zr[j + 1], a4.y, da);
                                da = fmaf(zr[j + 2], a4.z, da); da = fmaf(zr[j + 3], a4.w, da);
                            }
                            const float dista = dist_f32(zzc, ee[ka], da);
                            if (dista < bv) { bv2 = bv; bv = dista; bi = ka; }
                            else bv2 = fminf(bv2, dista);
                        }
                        nnear += near_tie(bv, bv2) ? 1u : 0u;
                    }
                }
                p.idx[seg][((size_t)b * C + c) * HWT + hw] = (long long)bi;
                if (p.fused) {
                    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.q) + ((size_t)b * C * D + (size_t)c * D) * HWT + hw;
                    const uint8_t* erow = ecb + bi * 128;
                    const uint32_t kx = (uint32_t)(bi & 7) << 4;
                    float ls0 = 0.0f, ls1 = 0.0f;
#pragma unroll
                    for (int j = 0; j < D; j += 4) {
                        const float4 e4 = *reinterpret_cast<const float4*>(erow + (((j >> 2) << 4) ^ kx));
                        const float d0 = __fsub_rn(e4.x, zr[j]), d1 = __fsub_rn(e4.y, zr[j + 1]);
                        const float d2 = __fsub_rn(e4.z, zr[j + 2]), d3 = __fsub_rn(e4.w, zr[j + 3]);
                        out[(size_t)j * HWT] = __float2bfloat16_rn(__fadd_rn(zr[j], d0));  // z + (q - z), models/vq_vae.py:53
                        out[(size_t)(j + 1) * HWT] = __float2bfloat16_rn(__fadd_rn(zr[j + 1], d1));
                        out[(size_t)(j + 2) * HWT] = __float2bfloat16_rn(__fadd_rn(zr[j + 2], d2));
                        out[(size_t)(j + 3) * HWT] = __float2bfloat16_rn(__fadd_rn(zr[j + 3], d3));
                        ls0 = fmaf(d0, d0, ls0); ls1 = fmaf(d1, d1, ls1);
                        ls0 = fmaf(d2, d2, ls0); ls1 = fmaf(d3, d3, ls1);
                    }
                    const float ls = ls0 + ls1;
#pragma unroll
                    for (int i = 0; i < C; ++i) lsum[i] += (i == c) ? ls : 0.0f;
                }
            }
        }
    }
    if (p.neartie && warp < 4 * NWG) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, nnear);
        if (lane == 0 && tot) atomicAdd(p.neartie, (unsigned long long)tot);
    }
    if (p.fused) {
        if (warp < 4 * NWG) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                double vv = (double)lsum[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vv += __shfl_xor_sync(0xffffffffu, vv, o);
                if (lane == 0 && vv != 0.0) atomicAdd(&p.loss_acc[c], vv);
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) *s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1u);
        __syncthreads();
        if (*s_last && tid == 0) {
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
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// 3-D tensor maps over the bf16 NCHW latents [B][Dtot][HW], box = 64 rows x `used` channels, SWIZZLE_128B
int make_maps_bf16(const QuantParams& p0, Maps& maps, int used) {
    if (!encode_fn()) return CTVQ_E_UNSUPPORTED;
    for (int sg = 0; sg < p0.n_seg; ++sg) {
        const cuuint64_t dims[3] = {(cuuint64_t)p0.HW, (cuuint64_t)p0.Dtot, (cuuint64_t)p0.B};
        const cuuint64_t strides[2] = {(cuuint64_t)p0.HW * 2, (cuuint64_t)p0.HW * p0.Dtot * 2};
        const cuuint32_t box[3] = {64u, (cuuint32_t)used, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = encode_fn()(&maps.m[sg], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<float*>(p0.z[sg]), dims, strides,
                                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return CTVQ_E_UNSUPPORTED;
    }
    return CTVQ_OK;
}

template <int D, int NK, int HWT, int C, int CS, int NSTAGE, int NWG>
int launch_bf16(const QuantParams& p0, cudaStream_t s) {
    Bf16Params P;
    P.q = p0;
    P.q.tiles_per_seg = (int)((p0.N + kTM - 1) / kTM);
    P.ntiles = P.q.tiles_per_seg * p0.n_seg;
    constexpr int USED = (C - 1) * CS + D, XR = (USED + 7) / 8 * 8, KP = (XR + 4 + 15) / 16 * 16;
    Maps maps;
    if (make_maps_bf16(p0, maps, USED) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    constexpr size_t smem = (size_t)NSTAGE * 2 * KP * 128 + (size_t)C * NK * 128 + (size_t)C * NK * 128 +
                            sizeof(float) * ((size_t)C * NK + ((C + 3) & ~3)) + (2 * NSTAGE + 4) * 8 + 16 + 1024;
    static_assert(smem <= 227 * 1024, "one CTA per SM");
    auto kern = vq_fwd_tc_bf16_kernel<D, NK, HWT, C, CS, NSTAGE, NWG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int grid = sm_count();
    if (grid > P.ntiles) grid = P.ntiles;
    kern<<<grid, 128 * NWG + 32, smem, s>>>(P, maps);
    return (int)cudaGetLastError();
}

}  // namespace

int launch_forward_tc_bf16(const QuantParams& p, cudaStream_t s) {
    if (p.dtype != CTVQ_BF16 || p.HW % 64 != 0 || p.K > 64 || p.N % 64 != 0) return CTVQ_E_UNSUPPORTED;
    for (int sg = 0; sg < p.n_seg; ++sg)
        if (reinterpret_cast<uintptr_t>(p.z[sg]) & 15) return CTVQ_E_UNSUPPORTED;
    for (int c = 0; c < p.C; ++c)
        if (reinterpret_cast<uintptr_t>(p.E[c]) & 15) return CTVQ_E_UNSUPPORTED;
    if (!encode_fn()) return CTVQ_E_UNSUPPORTED;
    if (p.d == 32 && p.cs == 1 && p.C == 4 && p.HW == 64) return launch_bf16<32, 64, 64, 4, 1, 6, 4>(p, s);
    if (p.d == 32 && p.cs == 1 && p.C == 4 && p.HW == 256) return launch_bf16<32, 64, 256, 4, 1, 6, 4>(p, s);
    if (p.d == 32 && p.cs == 1 && p.C == 2 && p.HW == 64) return launch_bf16<32, 64, 64, 2, 1, 6, 4>(p, s);
    return CTVQ_E_UNSUPPORTED;
}

}  // namespace ctvq
