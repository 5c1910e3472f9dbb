// Single-codebook tcgen05 forward kernel (VectorQuantizer of configs/vq_vae.yaml, the C=1 quantiser of
// configs/ct_mcq_vae.yaml, and the config-4 sweep shapes): same algorithm as ctvq_tc_fast.cu, but the two epilogue
// warpgroups split the ROWS instead of the codebooks — a CTA tile is 256 latent rows = two UMMA M-tiles, warpgroup g
// owns M-tile g and its own TMEM column range, so no cross-warpgroup combine is ever needed.  Codebooks wider than 256
// codes run as 256-column rounds over the same TMEM columns with an exact running best (every survivor of a round is
// re-scored with the exact fp32 formula, so rounds compare exact values; first index wins).
// Replaces models/vq_vae.py:25-55 and models/mcq_vae.py:26-74 for C = 1.
#include "ctvq_tc_ptx.cuh"

namespace ctvq {
using namespace tc;
namespace {

constexpr int kCT = 256;   // threads: 2 warpgroups x 4 warps
constexpr int kTR = 256;   // rows per CTA tile (2 M-tiles)

__device__ __forceinline__ void tmem_ld64c(uint32_t addr, float (&v)[64]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void or_if_le_c(unsigned& m, float a, float lim, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(m) : "f"(a), "f"(lim), "r"(bit));
}

struct C1Params {
    QuantParams q;
    int ntiles;
};

// D: channels; NK: codes padded to a multiple of 64 (and to a multiple of 256 when > 256); HWT: H*W; NSTAGE: TMA ring.
template <int D, int NK, int HWT, int NSTAGE, int MINB>
__global__ void __launch_bounds__(kCT, MINB) vq_fwd_tc_c1_kernel(const C1Params P, const __grid_constant__ Maps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const QuantParams& p = P.q;
    const int K = p.K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quarter = warp & 3, wg = warp >> 2;
    constexpr int NKR = NK > 256 ? 256 : NK;         // columns per round (per M-tile)
    constexpr int NROUND = NK / NKR;
    constexpr int NCH = NKR / 64;                    // 64-column register chunks per round
    constexpr int DJB = (D + 31) / 32;
    constexpr uint32_t kBlk = (uint32_t)D * 128u;    // one 32-row block: [D][128 B]
    constexpr uint32_t kStage = 8u * kBlk;           // 256 rows
    constexpr uint32_t kE = (uint32_t)DJB * NK * 128u;
    constexpr uint32_t kTmem = 2 * NKR <= 32 ? 32 : 2 * NKR <= 64 ? 64 : 2 * NKR <= 128 ? 128 : 2 * NKR <= 256 ? 256 : 512;
    static_assert(NK % 64 == 0 && NK % NKR == 0 && D % 8 == 0, "shape");
    uint8_t* a_s = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* e_s = a_s + (size_t)NSTAGE * kStage;
    float* ee_s = reinterpret_cast<float*>(e_s + kE);  // [NK]
    float* emax_s = ee_s + NK;                          // [4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(emax_s + 4);  // full[NSTAGE], mma
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NSTAGE + 1);
    const uint32_t a_base = smem_u32(a_s), e_base = smem_u32(e_s);
    const uint32_t bar_full0 = smem_u32(&bars[0]), bar_m = smem_u32(&bars[NSTAGE]);

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) mbar_init(bar_full0 + 8 * i, 1);
        mbar_init(bar_m, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmem);
    __syncthreads();

    const int niter = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto issue = [&](int it) {  // thread 0: TMA-load the 8 row blocks of iteration `it`
        const int tile = blockIdx.x + it * gridDim.x;
        const int seg = tile / p.tiles_per_seg;
        const long long row0 = (long long)(tile - seg * p.tiles_per_seg) * kTR;
        const int st = it % NSTAGE;
        int nblk = 0;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb) nblk += (row0 + 32 * mb < p.N) ? 1 : 0;
        mbar_expect_tx(bar_full0 + 8 * st, (uint32_t)nblk * kBlk);
        for (int mb = 0; mb < nblk; ++mb) {
            const long long nb = row0 + 32 * mb;
            const long long bb = nb / HWT;
            tma_load_3d(a_base + st * kStage + mb * kBlk, &maps.m[seg], bar_full0 + 8 * st, (int)(nb - bb * HWT), 0, (int)bb);
        }
    };
    if (tid == 0)
        for (int it = 0; it < NSTAGE - 1 && it < niter; ++it) issue(it);

    // ---- codebook -> K-major SWIZZLE_128B tile (once per persistent CTA) + |e|^2 ---------------------------------
    for (int i = tid; i < NK * DJB * 32; i += kCT) {
        const int j = i % (DJB * 32), k = i / (DJB * 32);
        const float v = (k < K && j < D) ? __ldg(p.E[0] + (size_t)k * D + j) : 0.0f;
        *reinterpret_cast<float*>(e_s + e_off(k, j, NK)) = v;
    }
    for (int k = tid; k < NK; k += kCT) {
        float a = CUDART_INF_F;
        if (k < K) {
            a = 0.0f;
            const float* row = p.E[0] + (size_t)k * D;
#pragma unroll 8
            for (int j = 0; j < D; ++j) { const float v = __ldg(row + j); a = fmaf(v, v, a); }
        }
        ee_s[k] = a;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {  // max |e_k|^2 over the real codes
        float mx = 0.0f;
        bool poisoned = false;  // fmaxf drops NaN: a NaN code norm must poison the bound (rows then take the exact scan)
        for (int k = lane; k < K; k += 32) { const float v = ee_s[k]; poisoned |= (v != v); mx = fmaxf(mx, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        poisoned = __any_sync(0xffffffffu, poisoned);
        if (lane == 0) emax_s[0] = poisoned ? CUDART_NAN_F : sqrtf(mx) * 1.0001f;
    }
    const uint32_t tmem_base = *tmem_slot;
    __syncthreads();
    const float emax = emax_s[0];

    uint32_t zsw[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) zsw[x] = ((((lane >> 3) ^ x) & 3) << 5) + ((lane & 7) << 2);

    uint32_t phase_m = 0;
    float lsum = 0.0f;
    unsigned nnear = 0u;  // near-tie rows seen by this thread (include/ctvq.h)

    for (int it = 0; it < niter; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int seg = tile / p.tiles_per_seg;
        const long long row0 = (long long)(tile - seg * p.tiles_per_seg) * kTR;
        const long long n = row0 + wg * 128 + quarter * 32 + lane;
        const bool valid = n < p.N;  // warp-uniform (N is a multiple of 32)
        const long long b = n / HWT;
        const int hw = (int)(n - b * HWT);
        const int st = it % NSTAGE;
        if (tid == 0) {
            if (NSTAGE == 1) issue(it);
            else if (it + NSTAGE - 1 < niter) issue(it + NSTAGE - 1);
        }
        mbar_wait_fast(bar_full0 + 8 * st, (uint32_t)(it / NSTAGE) & 1u);
        const uint32_t stage_u32 = a_base + st * kStage;
        const uint8_t* zrow = a_s + st * kStage + (wg * 4 + quarter) * kBlk;  // this thread's row block: [D][128 B]
        const int mtiles = (row0 + 128 < p.N) ? 2 : 1;

        float zz = 0.0f;
        float run_mn = CUDART_INF_F, run_bv = CUDART_INF_F, run_bv2 = CUDART_INF_F;  // exact best / second-best distance
        int run_bi = 0x7fffffff;
        bool run_bad = false;
#pragma unroll 1
        for (int rd = 0; rd < NROUND; ++rd) {
            tc_fence_after();
            if (tid == 0) {
                const uint32_t idesc = instr_desc_tf32(NKR);
                for (int mt = 0; mt < mtiles; ++mt) {
#pragma unroll
                    for (int s = 0; s < D / 8; ++s) {
                        const uint64_t ad = smem_desc(stage_u32 + mt * 4 * kBlk + s * 1024u, kBlk, 512u, 1u);
                        const uint64_t bd = smem_desc(e_base + (s >> 2) * NK * 128u + rd * NKR * 128u + (s & 3) * 32u, 16u, 1024u, 2u);
                        umma_tf32(tmem_base + mt * NKR, ad, bd, idesc, s > 0 ? 1u : 0u);
                    }
                }
                umma_commit(bar_m);
            }
            if (rd == 0 && valid) {  // |z|^2 (exact sequential chain) while the tensor core works
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float v = *reinterpret_cast<const float*>(zrow + j * 128 + zsw[j & 3]);
                    zz = fmaf(v, v, zz);
                }
            }
            mbar_wait_fast(bar_m, phase_m);
            phase_m ^= 1;
            tc_fence_after();
            if (valid) {
                const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + wg * NKR;
                const float* ee = ee_s + rd * NKR;
                float a[64];
                float m0 = CUDART_INF_F, m1 = CUDART_INF_F, m2 = CUDART_INF_F, m3 = CUDART_INF_F;
#pragma unroll
                for (int chn = 0; chn < NCH; ++chn) {
                    tmem_ld64c(trow + chn * 64, a);
#pragma unroll
                    for (int i = 0; i < 64; i += 4) {
                        const float4 e4 = *reinterpret_cast<const float4*>(ee + chn * 64 + i);
                        a[i] = fmaf(kNeg2OverC, a[i], e4.x); a[i + 1] = fmaf(kNeg2OverC, a[i + 1], e4.y);
                        a[i + 2] = fmaf(kNeg2OverC, a[i + 2], e4.z); a[i + 3] = fmaf(kNeg2OverC, a[i + 3], e4.w);
                        m0 = fminf(m0, a[i]); m1 = fminf(m1, a[i + 1]); m2 = fminf(m2, a[i + 2]); m3 = fminf(m3, a[i + 3]);
                    }
                }
                const float mn = fminf(fminf(m0, m1), fminf(m2, m3));
                run_mn = fminf(run_mn, mn);
                const float thr = 2.0f * (2.0f * kTf32Eps * sqrtf(zz) * 1.0001f * emax + kWinAbs * (zz + emax * emax));
                const float lim = run_mn + thr;  // running minimum: a superset of the final survivor set
                unsigned mask[NCH * 2];
                int cnt = 0;
#pragma unroll
                for (int chn = 0; chn < NCH; ++chn) {
                    if (NCH > 1) {
                        tmem_ld64c(trow + chn * 64, a);
#pragma unroll
                        for (int i = 0; i < 64; i += 4) {
                            const float4 e4 = *reinterpret_cast<const float4*>(ee + chn * 64 + i);
                            a[i] = fmaf(kNeg2OverC, a[i], e4.x); a[i + 1] = fmaf(kNeg2OverC, a[i + 1], e4.y);
                            a[i + 2] = fmaf(kNeg2OverC, a[i + 2], e4.z); a[i + 3] = fmaf(kNeg2OverC, a[i + 3], e4.w);
                        }
                    }
                    unsigned lo = 0u, hi = 0u;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        or_if_le_c(lo, a[i], lim, 1u << i);
                        or_if_le_c(hi, a[32 + i], lim, 1u << i);
                    }
                    mask[2 * chn] = lo;
                    mask[2 * chn + 1] = hi;
                    cnt += __popc(lo) + __popc(hi);
                }
                const bool finite = (zz < CUDART_INF_F) && (mn > -CUDART_INF_F) && (mn < CUDART_INF_F);
                if (!finite) run_bad = true;
                if (NROUND == 1 && finite && cnt == 1) {
#pragma unroll
                    for (int w = 0; w < NCH * 2; ++w)
                        if (mask[w]) run_bi = w * 32 + __ffs(mask[w]) - 1;
                } else if (finite) {
#pragma unroll
                    for (int w = 0; w < NCH * 2; ++w) {
                        unsigned mk = mask[w];
                        while (mk) {
                            const int k = rd * NKR + w * 32 + __ffs(mk) - 1;
                            mk &= mk - 1;
                            const uint8_t* erow = e_s + k * 128;
                            const uint32_t kx = (uint32_t)(k & 7) << 4;
                            float dot = 0.0f;
#pragma unroll
                            for (int j = 0; j < D; j += 4) {
                                const float4 e4 = *reinterpret_cast<const float4*>(erow + (j >> 5) * NK * 128 + ((((j & 31) >> 2) << 4) ^ kx));
                                dot = fmaf(*reinterpret_cast<const float*>(zrow + j * 128 + zsw[j & 3]), e4.x, dot);
                                dot = fmaf(*reinterpret_cast<const float*>(zrow + (j + 1) * 128 + zsw[(j + 1) & 3]), e4.y, dot);
                                dot = fmaf(*reinterpret_cast<const float*>(zrow + (j + 2) * 128 + zsw[(j + 2) & 3]), e4.z, dot);
                                dot = fmaf(*reinterpret_cast<const float*>(zrow + (j + 3) * 128 + zsw[(j + 3) & 3]), e4.w, dot);
                            }
                            const float dist = dist_f32(zz, ee_s[k], dot);
                            if (dist < run_bv) { run_bv2 = run_bv; run_bv = dist; run_bi = k; }  // ascending k: strict '<' keeps the first minimum
                            else run_bv2 = fminf(run_bv2, dist);
                        }
                    }
                }
            }
            if (rd + 1 < NROUND) {  // the next round overwrites the same TMEM columns
                tc_fence_before();
                __syncthreads();
            }
        }
        if (valid) {
            if (!run_bad && run_bi != 0x7fffffff) nnear += near_tie(run_bv, run_bv2) ? 1u : 0u;
            if (run_bad || run_bi == 0x7fffffff) {
                // non-finite row: exact scan of every code with torch.argmin's NaN rule
                run_bv = CUDART_INF_F; run_bi = 0x7fffffff;
                for (int k = 0; k < K; ++k) {
                    const uint8_t* erow = e_s + k * 128;
                    float dot = 0.0f;
#pragma unroll 8
                    for (int j = 0; j < D; ++j)
                        dot = fmaf(*reinterpret_cast<const float*>(zrow + j * 128 + zsw[j & 3]),
                                   *reinterpret_cast<const float*>(erow + (j >> 5) * NK * 128 + (((((j & 31) >> 2) ^ (k & 7)) & 7) << 4) + ((j & 3) << 2)), dot);
                    const float dist = dist_f32(zz, ee_s[k], dot);
                    if (k == 0 || (!(dist >= run_bv) && (run_bv == run_bv))) { run_bv = dist; run_bi = k; }  // k == 0 seeds the scan (all-+inf row -> 0)
                }
            }
            const int bi = run_bi;
            p.idx[seg][(size_t)b * HWT + hw] = (long long)bi;
            if (p.fused) {
                float* out = p.q + (size_t)b * D * HWT + hw;
                const uint8_t* erow = e_s + bi * 128;
                const uint32_t kx = (uint32_t)(bi & 7) << 4;
                float ls0 = 0.0f, ls1 = 0.0f;
#pragma unroll
                for (int j = 0; j < D; j += 4) {
                    const float4 e4 = *reinterpret_cast<const float4*>(erow + (j >> 5) * NK * 128 + ((((j & 31) >> 2) << 4) ^ kx));
                    const float z0 = *reinterpret_cast<const float*>(zrow + j * 128 + zsw[j & 3]);
                    const float z1 = *reinterpret_cast<const float*>(zrow + (j + 1) * 128 + zsw[(j + 1) & 3]);
                    const float z2 = *reinterpret_cast<const float*>(zrow + (j + 2) * 128 + zsw[(j + 2) & 3]);
                    const float z3 = *reinterpret_cast<const float*>(zrow + (j + 3) * 128 + zsw[(j + 3) & 3]);
                    const float d0 = __fsub_rn(e4.x, z0), d1 = __fsub_rn(e4.y, z1);
                    const float d2 = __fsub_rn(e4.z, z2), d3 = __fsub_rn(e4.w, z3);
                    out[(size_t)j * HWT] = __fadd_rn(z0, d0);
                    out[(size_t)(j + 1) * HWT] = __fadd_rn(z1, d1);
                    out[(size_t)(j + 2) * HWT] = __fadd_rn(z2, d2);
                    out[(size_t)(j + 3) * HWT] = __fadd_rn(z3, d3);
                    ls0 = fmaf(d0, d0, ls0); ls1 = fmaf(d1, d1, ls1);
                    ls0 = fmaf(d2, d2, ls0); ls1 = fmaf(d3, d3, ls1);
                }
                lsum += ls0 + ls1;
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM columns and this ring slot are free again
    }
    if (p.neartie) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, nnear);
        if (lane == 0 && tot) atomicAdd(p.neartie, (unsigned long long)tot);
    }
    if (p.fused) {
        double v = (double)lsum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(&p.loss_acc[0], v);
        __shared__ unsigned s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1u);
        __syncthreads();
        if (s_last && tid == 0) {
            __threadfence();
            const float m = (float)(__ldcg(&p.loss_acc[0]) / ((double)p.N * (double)D));
            const float l = __fadd_rn(__fmul_rn(m, p.beta), m);
            p.loss_out[0] = l;
            p.loss_out[1] = __fadd_rn(0.0f, l);
            p.loss_acc[0] = 0.0;
            *p.ticket = 0u;
            __threadfence();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmem);
}

template <int D, int NK, int HWT, int NSTAGE, int MINB>
int launch_c1(const QuantParams& p0, cudaStream_t s) {
    C1Params P;
    P.q = p0;
    P.q.tiles_per_seg = (int)((p0.N + kTR - 1) / kTR);
    P.ntiles = P.q.tiles_per_seg * p0.n_seg;
    constexpr int DJB = (D + 31) / 32;
    Maps maps;
    if (make_maps(p0, maps, D) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    constexpr size_t smem = (size_t)NSTAGE * 8 * D * 128 + (size_t)DJB * NK * 128 + sizeof(float) * (NK + 4) + (NSTAGE + 1) * 8 + 16 + 1024;
    static_assert(smem <= (MINB == 2 ? 113 : 225) * 1024, "shared memory budget");
    auto kern = vq_fwd_tc_c1_kernel<D, NK, HWT, NSTAGE, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int grid = sm_count() * MINB;
    if (grid > P.ntiles) grid = P.ntiles;
    kern<<<grid, kCT, smem, s>>>(P, maps);
    return (int)cudaGetLastError();
}
}  // namespace

int launch_forward_tc_c1(const QuantParams& p, cudaStream_t s) {
    if (p.C != 1 || p.HW % 32 != 0 || p.d != p.Dtot) return CTVQ_E_UNSUPPORTED;
    for (int sg = 0; sg < p.n_seg; ++sg)
        if (reinterpret_cast<uintptr_t>(p.z[sg]) & 15) return CTVQ_E_UNSUPPORTED;
    if (!encode_fn()) return CTVQ_E_UNSUPPORTED;
    // configs/vq_vae.yaml: K=512, D=64, latents [B,64,16,16]
    // (K in (256, 512] now takes the streaming kernel, ctvq_tc_stream.cu: 0.46 ms vs 0.54 ms at 1 M rows, 0.037 vs 0.059 ms at 16 K)
    // configs/ct_mcq_vae.yaml: K=64, d=128, latents [B,128,8,8]
    if (p.d == 128 && p.HW == 64 && p.K <= 64) return launch_c1<128, 64, 64, 1, 1>(p, s);
    // config-4 sweep shapes, HW = 256
    if (p.HW == 256 && p.K > 64 && p.K <= 256) {
        // (d = 32 now takes the streaming kernel: 0.162 ms vs 0.188 ms at 1 M rows)
        if (p.d == 64) return launch_c1<64, 256, 256, 2, 1>(p, s);
    }
    return CTVQ_E_UNSUPPORTED;
}

}  // namespace ctvq
