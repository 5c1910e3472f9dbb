// Streaming single-codebook tcgen05 forward kernel: ANY number of codes (the codebook streams through a TMA ring, it
// never has to fit in shared memory), D = 32 / 64 / 128 / 256 channels.  Covers VectorQuantizer of configs/vq_vae.yaml
// (K=512, D=64), the C=1 quantiser of configs/ct_mcq_vae.yaml (K=64, d=128) and the config-4 sweep up to K=16384.
// Replaces models/vq_vae.py:25-55 and models/mcq_vae.py:26-74 for C = 1.
//
// One persistent CTA per SM works on a SUPER-TILE of T x 128 latent rows (T "teams" of 4 epilogue warps, one UMMA
// M-tile each) whose NCHW slab [D channels x 128 rows] per team stays resident in shared memory while the codebook
// streams past it in 64-code UNITS:
//   TMA-A lane    super-tile slabs (cp.async.bulk.tensor.3d, SWIZZLE_128B_ATOM_32B = MN-major UMMA operand)
//   TMA-B lane    codebook blocks [64 codes x 32 channels] (2-D tensor map over the nn.Parameter, SWIZZLE_128B =
//                 K-major UMMA operand, rows >= K zero-filled by the TMA unit), ring of 8; every block is multiplied
//                 against ALL T teams before it is released, so one L2 read serves T x 128 rows;
//                 + the "|e|^2 block" of every 4 units (see below)
//   MMA lane      tcgen05.mma.kind::tf32 M=128 N=64 K=8, accumulating a unit's scores in one of the team's TMEM slots
//                 (512 columns = T teams x 8/T slots x 64 columns: units run up to 8/T - 1 ahead of the epilogue)
//   epilogue      thread = row: per unit pass 1 (max of s_k = z.e_k - |e_k|^2/2, which comes straight out of TMEM because
//                 |e_k|^2 rides in the GEMM: A = constant ones, B = -|e_k|^2/2 as three tf32 terms from a pre-pass),
//                 pass 2 (survivors within the rigorous tf32 bound of the RUNNING maximum), exact fp32 re-scoring of the
//                 survivors against the codebook in L2 -> exact running best (ascending k, first minimum wins);
//                 then gather / straight-through / loss from the winning row.
// Hand-over is mbarrier-only (no CTA-wide barrier in the loop).  The arithmetic contract is the same as
// every other kernel's, so indices equal the C oracle's on every row.
#include "ctvq_tc_ptx.cuh"

namespace ctvq {
using namespace tc;
namespace {

constexpr int kNSTB = 8;           // codebook-block ring depth
constexpr uint32_t kBlkB = 8192u;  // [64 codes][32 floats]

struct StreamParams {
    QuantParams q;
    int nsuper;          // super-tiles over all segments
    int NU;              // 64-code units (K padded to a multiple of 64)
    const float* ee;     // [NU*64] exact |e_k|^2 (+inf for padded codes)
    const unsigned* emax_bits;  // max |e_k|^2 as float bits
};
struct BMaps {
    CUtensorMap e;  // codebook [K][D], box 32 x 64
    CUtensorMap x;  // |e|^2 blocks [G*64][32], box 32 x 64
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void or_if_ge_s(unsigned& m, float a, float lim, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(m) : "f"(a), "f"(lim), "r"(bit));
}
__device__ __forceinline__ float sqrt_approx_s(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Pre-pass, one thread per (padded) code: exact |e_k|^2 (sequential chain of the arithmetic contract), its three-term
// tf32 split laid out as the B operand of the extra K-group, and the maximum.
__global__ void stream_prep_kernel(const float* __restrict__ E, int K, int D, int Kpad, float* __restrict__ ee,
                                   float* __restrict__ X, unsigned* __restrict__ emax_bits) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kpad) return;
    float a = CUDART_INF_F;
    if (k < K) {
        a = 0.0f;
        const float4* row = reinterpret_cast<const float4*>(E + (size_t)k * D);
        for (int m = 0; m < D / 4; ++m) {
            const float4 v = __ldg(row + m);
            a = fmaf(v.x, v.x, a); a = fmaf(v.y, v.y, a); a = fmaf(v.z, v.z, a); a = fmaf(v.w, v.w, a);
        }
        atomicMax(emax_bits, __float_as_uint(a));  // a >= 0 (or NaN / inf, which must poison the bound)
    }
    ee[k] = a;
    float t0 = -1.0e30f, t1 = 0.0f, t2 = 0.0f;  // padded / overflowed codes never survive the filter
    if (a < CUDART_INF_F) {
        const float h = (-0.5f * kTruncC) * a;  // centred truncation error (ctvq_common.cuh)
        t0 = __uint_as_float(__float_as_uint(h) & 0xFFFFE000u);
        const float r1 = h - t0;
        t1 = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
        t2 = __uint_as_float(__float_as_uint(r1 - t1) & 0xFFFFE000u);
    }
    float4* x = reinterpret_cast<float4*>(X + ((size_t)((k >> 8) * 64 + (k & 63))) * 32 + 8 * ((k >> 6) & 3));
    x[0] = make_float4(t0, t1, t2, 0.0f);
    x[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// D: channels; T: teams (UMMA M-tiles) per super-tile; ADB: slab buffers (2 = the next super-tile loads under this one)
template <int D, int T, int ADB>
__global__ void __launch_bounds__(128 * T + 96, 1) vq_fwd_tc_stream_kernel(const StreamParams P, const __grid_constant__ Maps maps,
                                                                            const __grid_constant__ BMaps bmaps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const QuantParams& p = P.q;
    const int K = p.K, HW = p.HW, NU = P.NU;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quarter = warp & 3, team = warp >> 2;
    constexpr int DJB = D / 32;
    constexpr int NBUF = 8 / T;                      // TMEM unit slots per team
    constexpr int TCOLS = 512 / T;
    constexpr bool ZREG = (D * T <= 128);            // row kept in registers when the register file allows (else re-read from the resident slab)
    constexpr uint32_t kBlkA = (uint32_t)D * 128u;   // one 32-row block: [D][128 B]
    constexpr uint32_t kASZ = (uint32_t)T * 4u * kBlkA;
    static_assert(D % 32 == 0 && (T == 1 || T == 2 || T == 4), "shape");
    uint8_t* a_s = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_s = a_s + (size_t)ADB * kASZ;
    uint8_t* x_s = b_s + (size_t)kNSTB * kBlkB;
    uint8_t* ones_s = x_s + 2 * kBlkB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones_s + 4096);
    // afull[ADB] aempty[ADB] bfull[8] bempty[8] xfull[2] xempty[2] mma[T*NBUF=8] tfree[8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ADB + 2 * kNSTB + 4 + 16);
    const uint32_t a_base = smem_u32(a_s), b_base = smem_u32(b_s), x_base = smem_u32(x_s), ones_base = smem_u32(ones_s);
    const uint32_t bar_afull = smem_u32(&bars[0]), bar_aempty = bar_afull + 8 * ADB;
    const uint32_t bar_bfull = bar_aempty + 8 * ADB, bar_bempty = bar_bfull + 8 * kNSTB;
    const uint32_t bar_xfull = bar_bempty + 8 * kNSTB, bar_xempty = bar_xfull + 16;
    const uint32_t bar_mma = bar_xempty + 16, bar_tfree = bar_mma + 64;

    if (tid == 0) {
        for (int i = 0; i < ADB; ++i) { mbar_init(bar_afull + 8 * i, 1); mbar_init(bar_aempty + 8 * i, 1 + 4 * T); }
        for (int i = 0; i < kNSTB; ++i) { mbar_init(bar_bfull + 8 * i, 1); mbar_init(bar_bempty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_xfull + 8 * i, 1); mbar_init(bar_xempty + 8 * i, 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(bar_mma + 8 * i, 1); mbar_init(bar_tfree + 8 * i, 4); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    for (int i = tid; i < 1024; i += 128 * T + 96)  // rows are constant, so the swizzle inside a 128-byte row is immaterial
        reinterpret_cast<float*>(ones_s)[i] = ((i >> 5) & 7) < 3 ? 1.0f : 0.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nsup = (P.nsuper - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // super-tiles of this CTA
    const int NXG = (NU + 3) >> 2;
    float lsum = 0.0f;
    unsigned nnear = 0u;  // near-tie rows seen by this thread (include/ctvq.h)

    if (warp == 4 * T + 2) {
        // ===================================== TMA-A lane: super-tile slabs ==========================================
        if (lane == 0) {
            for (int s = 0; s < nsup; ++s) {
                const int ab = s % ADB;
                if (s >= ADB) mbar_wait_fast(bar_aempty + 8 * ab, (uint32_t)(s / ADB - 1) & 1u);
                const int sup = blockIdx.x + s * gridDim.x;
                const int seg = sup / p.tiles_per_seg;
                const long long row0 = (long long)(sup - seg * p.tiles_per_seg) * (128 * T);
                int nblk = 0;
                for (int mb = 0; mb < 4 * T; ++mb) nblk += (row0 + 32 * mb < p.N) ? 1 : 0;
                mbar_expect_tx(bar_afull + 8 * ab, (uint32_t)nblk * kBlkA);
                for (int mb = 0; mb < nblk; ++mb) {
                    const long long nb = row0 + 32 * mb;
                    const long long bb = nb / HW;
                    tma_load_3d(a_base + ab * kASZ + mb * kBlkA, &maps.m[seg], bar_afull + 8 * ab, (int)(nb - bb * HW), 0, (int)bb);
                }
            }
        }
    } else if (warp == 4 * T + 1) {
        // ===================================== TMA-B lane: codebook blocks + |e|^2 blocks ============================
        if (lane == 0) {
            long long bbk = 0;  // running block counter
            int xc = 0;         // running |e|^2-block counter
            for (int s = 0; s < nsup; ++s) {
                for (int kc = 0; kc < NU; ++kc) {
                    if ((kc & 3) == 0) {
                        const int xs = xc & 1;
                        if (xc >= 2) mbar_wait_fast(bar_xempty + 8 * xs, (uint32_t)((xc >> 1) - 1) & 1u);
                        mbar_expect_tx(bar_xfull + 8 * xs, kBlkB);
                        tma_load_2d(x_base + xs * kBlkB, &bmaps.x, bar_xfull + 8 * xs, 0, (kc >> 2) * 64);
                        ++xc;
                    }
                    for (int db = 0; db < DJB; ++db, ++bbk) {
                        const int bs = (int)(bbk % kNSTB);
                        if (bbk >= kNSTB) mbar_wait_fast(bar_bempty + 8 * bs, (uint32_t)(bbk / kNSTB - 1) & 1u);
                        mbar_expect_tx(bar_bfull + 8 * bs, kBlkB);
                        tma_load_2d(b_base + bs * kBlkB, &bmaps.e, bar_bfull + 8 * bs, db * 32, kc * 64);
                    }
                }
            }
        }
    } else if (warp == 4 * T) {
        // ===================================== MMA lane ==============================================================
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32(64);
            long long bbk = 0;
            int xc = -1;
            for (int s = 0; s < nsup; ++s) {
                const int ab = s % ADB;
                mbar_wait_fast(bar_afull + 8 * ab, (uint32_t)(s / ADB) & 1u);
                const uint32_t slab = a_base + ab * kASZ;
                for (int kc = 0; kc < NU; ++kc) {
                    const long long g = (long long)s * NU + kc;
                    const int slot = (int)(g % NBUF);
                    if (g >= NBUF)
                        for (int t = 0; t < T; ++t) mbar_wait_fast(bar_tfree + 8 * (t * NBUF + slot), (uint32_t)(g / NBUF - 1) & 1u);
                    if ((kc & 3) == 0) {
                        ++xc;
                        mbar_wait_fast(bar_xfull + 8 * (xc & 1), (uint32_t)(xc >> 1) & 1u);
                    }
                    tc_fence_after();
                    for (int db = 0; db < DJB; ++db, ++bbk) {
                        const int bs = (int)(bbk % kNSTB);
                        mbar_wait_fast(bar_bfull + 8 * bs, (uint32_t)(bbk / kNSTB) & 1u);
                        tc_fence_after();
#pragma unroll
                        for (int t = 0; t < T; ++t) {
#pragma unroll
                            for (int s8 = 0; s8 < 4; ++s8) {
                                const uint64_t ad = smem_desc(slab + (uint32_t)t * 4u * kBlkA + (uint32_t)(db * 4 + s8) * 1024u, kBlkA, 512u, 1u);
                                const uint64_t bd = smem_desc(b_base + bs * kBlkB + s8 * 32u, 16u, 1024u, 2u);
                                umma_tf32(tmem_base + t * TCOLS + slot * 64, ad, bd, idesc, (db > 0 || s8 > 0) ? 1u : 0u);
                            }
                        }
                        umma_commit(bar_bempty + 8 * bs);
                    }
                    // + 1 * (-|e_k|^2 / 2): the accumulator now holds the whole score z.e_k - |e_k|^2/2
#pragma unroll
                    for (int t = 0; t < T; ++t)
                        umma_tf32(tmem_base + t * TCOLS + slot * 64, smem_desc(ones_base, 1024u, 512u, 1u),
                                  smem_desc(x_base + (xc & 1) * kBlkB + (kc & 3) * 32u, 16u, 1024u, 2u), idesc, 1u);
                    if ((kc & 3) == 3 || kc == NU - 1) umma_commit(bar_xempty + 8 * (xc & 1));
#pragma unroll
                    for (int t = 0; t < T; ++t) umma_commit(bar_mma + 8 * (t * NBUF + slot));
                }
                umma_commit(bar_aempty + 8 * ab);  // the tensor core is done with this super-tile's slabs
            }
        }
    } else {
        // ===================================== epilogue: thread = latent row =========================================
        uint32_t zsw[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) zsw[x] = ((((lane >> 3) ^ x) & 3) << 5) + ((lane & 7) << 2);
        const float* __restrict__ E = p.E[0];
        const float emax = sqrtf(__uint_as_float(__ldg(P.emax_bits))) * 1.0001f;
        for (int s = 0; s < nsup; ++s) {
            const int ab = s % ADB;
            const int sup = blockIdx.x + s * gridDim.x;
            const int seg = sup / p.tiles_per_seg;
            const long long row0 = (long long)(sup - seg * p.tiles_per_seg) * (128 * T);
            const long long n = row0 + team * 128 + quarter * 32 + lane;
            const bool valid = n < p.N;  // warp-uniform (N is a multiple of 32)
            const long long b = n / HW;
            const int hw = (int)(n - b * HW);
            mbar_wait_fast(bar_afull + 8 * ab, (uint32_t)(s / ADB) & 1u);
            const uint8_t* zrow = a_s + ab * kASZ + (team * 4 + quarter) * kBlkA;  // this thread's row block: [D][128 B]
            float zr[ZREG ? D : 1];
            float zz = 0.0f;  // exact sequential chain (arithmetic contract)
            if (valid) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float v = *reinterpret_cast<const float*>(zrow + j * 128 + zsw[j & 3]);
                    if (ZREG) zr[j] = v;
                    zz = fmaf(v, v, zz);
                }
            }
            if (ZREG) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_aempty + 8 * ab);  // the row lives in registers from here on
            }
            auto zat = [&](int j) -> float {
                return ZREG ? zr[ZREG ? j : 0] : *reinterpret_cast<const float*>(zrow + j * 128 + zsw[j & 3]);
            };
            // rigorous bound on |tf32 distance - exact-chain distance|; scores are distances / -2
            const float thr = 2.0f * (2.0f * kTf32Eps * sqrt_approx_s(zz) * 1.0001f * emax + kWinAbs * (zz + emax * emax));
            // Running state.  Exact distances are only needed to COMPARE candidates, and an L2 round trip per unit would stall
            // the few warps an SM has, so survivors are QUEUED (three register slots: code + an upper bound of its
            // approximate score = the running maximum when it was queued, hence non-decreasing along the queue).  A later
            // unit that lifts the window above the oldest bounds drops them unscored -- the usual fate of every provisional
            // maximum -- and what is left at the end of the row is scored in one go, two chains interleaved; a row whose
            // final window holds a single code never computes an exact distance.  Queue overflow settles it exactly.
            float run_mx = -CUDART_INF_F, run_bv = CUDART_INF_F, run_bv2 = CUDART_INF_F, ex_ub = -CUDART_INF_F;
            float qu0 = 0.0f, qu1 = 0.0f, qu2 = 0.0f;
            int run_bi = 0x7fffffff, qk0 = -1, qk1 = -1, qk2 = -1, qn = 0;
            bool bad = !(zz < CUDART_INF_F);
            auto score = [&](int k) {  // exact fp32 distance of code k (arithmetic contract); callers go in ascending k
                const float4* erow = reinterpret_cast<const float4*>(E + (size_t)k * D);
                float dot = 0.0f;
#pragma unroll(ZREG ? D / 4 : 8)
                for (int j = 0; j < D; j += 4) {
                    const float4 e4 = __ldg(erow + (j >> 2));
                    dot = fmaf(zat(j), e4.x, dot);
                    dot = fmaf(zat(j + 1), e4.y, dot);
                    dot = fmaf(zat(j + 2), e4.z, dot);
                    dot = fmaf(zat(j + 3), e4.w, dot);
                }
                const float dist = dist_f32(zz, __ldg(P.ee + k), dot);
                if (dist < run_bv) { run_bv2 = run_bv; run_bv = dist; run_bi = k; }  // strict '<' keeps the first minimum
                else run_bv2 = fminf(run_bv2, dist);
            };
            auto score2 = [&](int ka, int kb) {  // two codes (ka < kb, kb may be -1): both rows in flight, chains interleaved
                const float4* ra = reinterpret_cast<const float4*>(E + (size_t)ka * D);
                const float4* rb = reinterpret_cast<const float4*>(E + (size_t)(kb >= 0 ? kb : ka) * D);
                float da = 0.0f, db = 0.0f;
#pragma unroll(ZREG ? D / 4 : 8)
                for (int j = 0; j < D; j += 4) {
                    const float4 a4 = __ldg(ra + (j >> 2));
                    const float4 b4 = __ldg(rb + (j >> 2));
                    const float z0 = zat(j), z1 = zat(j + 1), z2 = zat(j + 2), z3 = zat(j + 3);
                    da = fmaf(z0, a4.x, da); db = fmaf(z0, b4.x, db);
                    da = fmaf(z1, a4.y, da); db = fmaf(z1, b4.y, db);
                    da = fmaf(z2, a4.z, da); db = fmaf(z2, b4.z, db);
                    da = fmaf(z3, a4.w, da); db = fmaf(z3, b4.w, db);
                }
                const float dista = dist_f32(zz, __ldg(P.ee + ka), da);
                if (dista < run_bv) { run_bv2 = run_bv; run_bv = dista; run_bi = ka; }
                else run_bv2 = fminf(run_bv2, dista);
                if (kb >= 0) {
                    const float distb = dist_f32(zz, __ldg(P.ee + kb), db);
                    if (distb < run_bv) { run_bv2 = run_bv; run_bv = distb; run_bi = kb; }
                    else run_bv2 = fminf(run_bv2, distb);
                }
            };
#pragma unroll 1
            for (int kc = 0; kc < NU; ++kc) {
                const long long g = (long long)s * NU + kc;
                const int slot = (int)(g % NBUF);
                mbar_wait_fast(bar_mma + 8 * (team * NBUF + slot), (uint32_t)(g / NBUF) & 1u);
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + team * TCOLS + slot * 64;
                unsigned mask0 = 0u, mask1 = 0u;
                float mxu = 0.0f;
                bool hit = false;  // some score of this unit lies inside this row's window
                if (valid) {
                    uint32_t a[32];
                    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        tmem_ld32_issue(trow + 32 * h, a);
                        tmem_ld32_wait(a);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            m0 = fmaxf(m0, __uint_as_float(a[i])); m1 = fmaxf(m1, __uint_as_float(a[i + 1]));
                            m2 = fmaxf(m2, __uint_as_float(a[i + 2])); m3 = fmaxf(m3, __uint_as_float(a[i + 3]));
                        }
                    }
                    mxu = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    run_mx = fmaxf(run_mx, mxu);
                    hit = mxu >= run_mx - 0.5f * thr;
                }
                // pass 2 only when SOME row of this warp has a score of this unit inside its window (a unit whose maximum is
                // below the row's limit has no survivor): with K in the thousands the running maximum settles early and most
                // later units are skipped -- a warp of 32 rows runs pass 2 on ~ 32 (1 + ln(units / 32)) units (98 of 256 at
                // K = 16384).  The vote is taken by the whole (converged) warp.
                if (__any_sync(0xffffffffu, hit)) {
                    uint32_t a[32];
                    const float lim = run_mx - 0.5f * thr;  // running maximum: a superset of the final survivor set
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        tmem_ld32_issue(trow + 32 * h, a);
                        tmem_ld32_wait(a);
                        unsigned mk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                        for (int i = 0; i < 32; ++i) or_if_ge_s(mk[i & 3], __uint_as_float(a[i]), lim, 1u << i);
                        const unsigned m = (mk[0] | mk[1]) | (mk[2] | mk[3]);
                        if (h == 0) mask0 = m; else mask1 = m;
                    }
                    if (!valid) { mask0 = 0u; mask1 = 0u; }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tfree + 8 * (team * NBUF + slot));  // slot free for unit g + NBUF
                if (valid) {
                    if (!(mxu > -CUDART_INF_F) || !(mxu < CUDART_INF_F)) bad = true;
                    if (!bad) {
                        unsigned long long mk = ((unsigned long long)mask1 << 32) | mask0;
                        const float lim = run_mx - 0.5f * thr;
                        // everything older whose approximate score cannot reach the window any more is out: a PREFIX of the queue
                        const int drop = ((qn > 0 && qu0 < lim) ? 1 : 0) + ((qn > 1 && qu1 < lim) ? 1 : 0) + ((qn > 2 && qu2 < lim) ? 1 : 0);
                        if (drop == 1) { qk0 = qk1; qu0 = qu1; qk1 = qk2; qu1 = qu2; }
                        else if (drop == 2) { qk0 = qk2; qu0 = qu2; }
                        qn -= drop;
                        if (run_bi != 0x7fffffff && ex_ub < lim) { run_bi = 0x7fffffff; run_bv = CUDART_INF_F; run_bv2 = CUDART_INF_F; }
                        while (mk) {
                            const int k = kc * 64 + __ffsll((long long)mk) - 1;
                            mk &= mk - 1;
                            if (k >= K) continue;
                            if (qn < 3) {
                                if (qn == 0) { qk0 = k; qu0 = run_mx; } else if (qn == 1) { qk1 = k; qu1 = run_mx; } else { qk2 = k; qu2 = run_mx; }
                                ++qn;
                            } else {  // overflow: settle the queue exactly (ascending k), then this code
                                score2(qk0, qk1);
                                score2(qk2, k);
                                qn = 0;
                                ex_ub = run_mx;
                            }
                        }
                    }
                }
            }
            if (valid && !bad) {
                if (qn == 1 && run_bi == 0x7fffffff) {
                    run_bi = qk0;  // the only code in the final window: no exact distance needed
                } else if (qn >= 1) {  // queued codes are younger (larger k) than every scored entry: ascending order holds
                    score2(qk0, qn >= 2 ? qk1 : -1);
                    if (qn == 3) score(qk2);
                }
            }
            if (valid && !bad && run_bi != 0x7fffffff) nnear += near_tie(run_bv, run_bv2) ? 1u : 0u;
            if (valid) {
                if (bad || run_bi == 0x7fffffff) {
                    // non-finite row: exact scan of every code with torch.argmin's NaN rule
                    run_bv = CUDART_INF_F; run_bi = 0x7fffffff;
                    for (int k = 0; k < K; ++k) {
                        const float* erow = E + (size_t)k * D;
                        float dot = 0.0f;
#pragma unroll(ZREG ? D : 8)
                        for (int j = 0; j < D; ++j) dot = fmaf(zat(j), __ldg(erow + j), dot);
                        const float dist = dist_f32(zz, __ldg(P.ee + k), dot);
                        if (k == 0 || (!(dist >= run_bv) && (run_bv == run_bv))) { run_bv = dist; run_bi = k; }  // k == 0 seeds the scan (all-+inf row -> 0)
                    }
                }
                const int bi = run_bi;
                p.idx[seg][(size_t)b * HW + hw] = (long long)bi;
                if (p.fused) {
                    float* out = p.q + (size_t)b * D * HW + hw;  // a warp's 32 rows are contiguous: 128-byte stores
                    const float4* erow = reinterpret_cast<const float4*>(E + (size_t)bi * D);
                    float ls0 = 0.0f, ls1 = 0.0f;
#pragma unroll(ZREG ? D / 4 : 8)
                    for (int j = 0; j < D; j += 4) {
                        const float4 e4 = __ldg(erow + (j >> 2));
                        const float z0 = zat(j), z1 = zat(j + 1), z2 = zat(j + 2), z3 = zat(j + 3);
                        const float d0 = __fsub_rn(e4.x, z0), d1 = __fsub_rn(e4.y, z1);
                        const float d2 = __fsub_rn(e4.z, z2), d3 = __fsub_rn(e4.w, z3);
                        out[(size_t)j * HW] = __fadd_rn(z0, d0);  // z + (q - z), models/vq_vae.py:53
                        out[(size_t)(j + 1) * HW] = __fadd_rn(z1, d1);
                        out[(size_t)(j + 2) * HW] = __fadd_rn(z2, d2);
                        out[(size_t)(j + 3) * HW] = __fadd_rn(z3, d3);
                        ls0 = fmaf(d0, d0, ls0); ls1 = fmaf(d1, d1, ls1);
                        ls0 = fmaf(d2, d2, ls0); ls1 = fmaf(d3, d3, ls1);
                    }
                    lsum += ls0 + ls1;
                }
            }
            if (!ZREG) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_aempty + 8 * ab);  // done reading the slab
            }
        }
    }
    if (p.neartie && warp < 4 * T) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, nnear);
        if (lane == 0 && tot) atomicAdd(p.neartie, (unsigned long long)tot);
    }
    // ---- loss: warp sums -> fp64 atomics -> last CTA finalises -----------------------------------------------------
    if (p.fused) {
        if (warp < 4 * T) {
            double v = (double)lsum;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(&p.loss_acc[0], v);
        }
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
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int make_map_2d(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows) {
    if (!encode_fn()) return CTVQ_E_UNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32u, 64u};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CTVQ_OK : CTVQ_E_UNSUPPORTED;
}

template <int D, int T, int ADB>
int launch_stream(const QuantParams& p0, cudaStream_t s) {
    const int NU = (p0.K + 63) / 64, Kpad = ((p0.K + 255) / 256) * 256, G = Kpad / 256;
    // scratch: [emax bits, pad to 256 B][ee: Kpad floats][X: G*64*32 floats]
    const size_t need = 256 + (size_t)Kpad * 4 + (size_t)G * 8192;
    if (!p0.scratch || p0.scratch_bytes < need || (reinterpret_cast<uintptr_t>(p0.scratch) & 255)) return CTVQ_E_UNSUPPORTED;
    unsigned* emax_bits = reinterpret_cast<unsigned*>(p0.scratch);
    float* ee = reinterpret_cast<float*>(p0.scratch + 256);
    float* X = ee + Kpad;
    StreamParams P;
    P.q = p0;
    P.q.tiles_per_seg = (int)((p0.N + 128 * T - 1) / (128 * T));
    P.nsuper = P.q.tiles_per_seg * p0.n_seg;
    P.NU = NU;
    P.ee = ee;
    P.emax_bits = emax_bits;
    Maps maps;
    if (make_maps(p0, maps, D) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    BMaps bm;
    if (make_map_2d(&bm.e, p0.E[0], (uint64_t)D, (uint64_t)p0.K) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    if (make_map_2d(&bm.x, X, 32, (uint64_t)G * 64) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    cudaError_t e = cudaMemsetAsync(emax_bits, 0, 4, s);
    if (e != cudaSuccess) return (int)e;
    stream_prep_kernel<<<Kpad / 256, 256, 0, s>>>(p0.E[0], p0.K, D, Kpad, ee, X, emax_bits);
    constexpr size_t smem = (size_t)ADB * T * 4 * D * 128 + (size_t)kNSTB * kBlkB + 2 * kBlkB + 4096 +
                            (2 * ADB + 2 * kNSTB + 4 + 16) * 8 + 16 + 1024;
    static_assert(smem <= 227 * 1024, "one CTA per SM");
    auto kern = vq_fwd_tc_stream_kernel<D, T, ADB>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int grid = sm_count();
    if (grid > P.nsuper) grid = P.nsuper;
    kern<<<grid, 128 * T + 96, smem, s>>>(P, maps, bm);
    return (int)cudaGetLastError();
}

}  // namespace

size_t stream_scratch_bytes(int K) {
    const size_t Kpad = ((size_t)(K > 0 ? K : 1) + 255) / 256 * 256;
    return 256 + Kpad * 4 + (Kpad / 256) * 8192;
}

bool stream_supported(const QuantParams& p) {
    if (p.C != 1 || p.HW % 32 != 0 || p.d != p.Dtot) return false;
    if (p.d != 32 && p.d != 64 && p.d != 128 && p.d != 256) return false;
    for (int sg = 0; sg < p.n_seg; ++sg)
        if (reinterpret_cast<uintptr_t>(p.z[sg]) & 15) return false;
    if (reinterpret_cast<uintptr_t>(p.E[0]) & 15) return false;
    if (!p.scratch || p.scratch_bytes < stream_scratch_bytes(p.K) || (reinterpret_cast<uintptr_t>(p.scratch) & 255)) return false;
    return encode_fn() != nullptr;
}

int launch_forward_tc_stream(const QuantParams& p, cudaStream_t s) {
    if (!stream_supported(p)) return CTVQ_E_UNSUPPORTED;
    if (p.d == 32) return launch_stream<32, 4, 2>(p, s);
    if (p.d == 64) return launch_stream<64, 2, 2>(p, s);  // rows in registers + double-buffered slabs: 0.46 ms vs 0.63 ms (T=4) at K=512, 1 M rows
    if (p.d == 128) return launch_stream<128, 2, 1>(p, s);
    if (p.d == 256) return launch_stream<256, 1, 1>(p, s);
    return CTVQ_E_UNSUPPORTED;
}

}  // namespace ctvq
