// tcgen05 / TMEM / TMA path of the quantiser forward (sm_100a).
//
// Replaces the distance GEMM + argmin (+ gather/loss) of models/vq_vae.py:30-55 and models/mcq_vae.py:26-64
// for all C codebooks of models/mcq_vae.py:100-127 in one launch:
//
//   TMA (cp.async.bulk.tensor.3d, SWIZZLE_128B) pulls, per codebook, a [d channels x 128 rows] slab of the NCHW
//   latents straight into the MN-major UMMA operand layout — no NCHW->NHWC transpose ever exists;
//   the codebooks sit in shared memory in the K-major SWIZZLE_128B layout for the whole (persistent) CTA;
//   one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=codes, K=8 per instruction), accumulating
//   z.e^T for up to 256 (codebook, code) columns in TMEM;
//   the epilogue reads the accumulators with tcgen05.ld (one latent row per thread), forms the approximate
//   distances |e|^2 - 2 z.e, and keeps every code within a rigorous error bound of the row minimum;
//   rows with more than one surviving code re-score the survivors with the EXACT fp32 sequential-FMA formula of
//   the arithmetic contract, so the indices are bit-identical to the SIMT kernel and the C oracle;
//   gather + straight-through + loss are fused behind it, reading z from the same shared-memory slab.
//
// The N x K distance matrix lives only in TMEM.
#include "ctvq_tc_ptx.cuh"

namespace ctvq {
using namespace tc;
namespace {

constexpr int kThreads = 128;   // 4 warps: warp w owns TMEM lanes 32w..32w+31
constexpr int kTmemCols = 256;  // accumulator columns per CTA (2 CTAs/SM share the 512)
constexpr int kMaxBlk = kTmemCols / 32;

struct TcParams {
    QuantParams q;
    int Kpad;       // K rounded up to a multiple of 16 (UMMA N granularity at M=128)
    int ntiles;     // over all segments
    int units;      // (codebook, column-chunk) pairs
    int chunks;     // column chunks per codebook (ceil(Kpad/256))
    int djb;        // 128-byte blocks per codebook row (ceil(d/32))
    unsigned a_bytes, e_bytes;  // shared-memory bytes of the A slabs / codebook tiles
    float* dbg;     // debug: raw TMEM dot products [ntiles*128][C*Kpad] (null in production)
};

__global__ void __launch_bounds__(kThreads) vq_fwd_tc_kernel(const TcParams P, const __grid_constant__ Maps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const QuantParams& p = P.q;
    const int C = p.C, d = p.d, K = p.K, HW = p.HW, Kpad = P.Kpad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- shared-memory carve-up (all bases 1024-byte aligned) --------------------------------------------
    uint8_t* a_s = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // C slabs x 4 row-blocks x [d][128B]
    uint8_t* e_s = a_s + P.a_bytes;              // C x djb x [Kpad][128B]
    float* ee_s = reinterpret_cast<float*>(e_s + P.e_bytes);  // [C][Kpad]
    float* emax_s = ee_s + C * Kpad;             // [C]
    float* lsum_s = emax_s + ((C + 1) & ~1);     // [C][128] thread-private loss partials
    double* red = reinterpret_cast<double*>(lsum_s + C * kThreads);  // [4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 4);  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const uint32_t a_base = smem_u32(a_s), e_base = smem_u32(e_s);
    const uint32_t bar_a = smem_u32(&bars[0]), bar_m = smem_u32(&bars[1]);
    const uint32_t slab_bytes = 4u * d * 128u;   // one codebook's A slab
    const uint32_t ecb_bytes = (uint32_t)P.djb * Kpad * 128u;

    if (tid == 0) {
        mbar_init(bar_a, 1);
        mbar_init(bar_m, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    // ---- stage the codebooks (generic proxy, swizzled by hand) + |e|^2 -------------------------------------
    {
        const int jw = P.djb * 32;
        for (int i = tid; i < C * Kpad * jw; i += kThreads) {
            const int j = i % jw, ck = i / jw, k = ck % Kpad, c = ck / Kpad;
            const float v = (k < K && j < d) ? __ldg(p.E[c] + (size_t)k * d + j) : 0.0f;
            *reinterpret_cast<float*>(e_s + (size_t)c * ecb_bytes + e_off(k, j, Kpad)) = v;
        }
        for (int i = tid; i < C * Kpad; i += kThreads) {
            const int k = i % Kpad, c = i / Kpad;
            float a = CUDART_INF_F;
            if (k < K) {
                a = 0.0f;
                const float* row = p.E[c] + (size_t)k * d;
                for (int j = 0; j < d; ++j) { const float v = __ldg(row + j); a = fmaf(v, v, a); }
            }
            ee_s[i] = a;
        }
    }
    fence_proxy_async();  // codebook tiles were written through the generic proxy; tcgen05.mma reads via the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid < C) {
        float mx = 0.0f;
        bool poisoned = false;  // fmaxf drops NaN: a NaN code norm must poison the bound (the row then takes the exact scan)
        for (int k = 0; k < K; ++k) { const float v = ee_s[tid * Kpad + k]; poisoned |= (v != v); mx = fmaxf(mx, v); }
        emax_s[tid] = poisoned ? CUDART_NAN_F : sqrtf(mx) * 1.0001f;
    }
    const uint32_t tmem_base = *tmem_slot;
    __syncthreads();

    uint32_t phase_a = 0, phase_m = 0;
    unsigned nnear = 0u;  // near-tie rows seen by this thread (include/ctvq.h)
    for (int c = 0; c < C; ++c) lsum_s[c * kThreads + tid] = 0.0f;

    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int seg = tile / p.tiles_per_seg;
        const long long row0 = (long long)(tile - seg * p.tiles_per_seg) * kTM;
        const long long n = row0 + tid;
        const bool valid = n < p.N;  // warp-uniform: N and row blocks are multiples of 32
        const long long b = n / HW;
        const int hw = (int)(n - b * HW);
        // ---- TMA: one [32 rows x d channels] box per (codebook, row block) -------------------------------------
        if (tid == 0) {
            int nblk = 0;
            for (int mb = 0; mb < 4; ++mb) nblk += (row0 + 32 * mb < p.N) ? 1 : 0;
            mbar_expect_tx(bar_a, (uint32_t)(C * nblk) * d * 128u);
            for (int c = 0; c < C; ++c)
                for (int mb = 0; mb < nblk; ++mb) {
                    const long long nb = row0 + 32 * mb;
                    const long long bb = nb / HW;
                    tma_load_3d(a_base + c * slab_bytes + mb * d * 128u, &maps.m[seg], bar_a, (int)(nb - bb * HW),
                                c * p.cs, (int)bb);
                }
        }
        mbar_wait(bar_a, phase_a);
        phase_a ^= 1;

        // exact running best across the column chunks of one codebook (only used when a codebook spans > 256 codes)
        float run_mn = CUDART_INF_F, run_bv = CUDART_INF_F, run_bv2 = CUDART_INF_F;  // exact best / second-best distance
        int run_bi = 0x7fffffff;
        bool run_bad = false;
        for (int u0 = 0; u0 < P.units;) {
            // ---- pack units (codebook, column chunk) into <= 256 TMEM columns --------------------------------------
            int u1 = u0, cols = 0;
            while (u1 < P.units) {
                const int ch = u1 % P.chunks;
                const int w = min(256, Kpad - ch * 256);
                if (cols + w > kTmemCols) break;
                cols += w;
                ++u1;
            }
            tc_fence_after();
            if (tid == 0) {
                int col = 0;
                for (int u = u0; u < u1; ++u) {
                    const int c = u / P.chunks, ch = u % P.chunks;
                    const int w = min(256, Kpad - ch * 256);
                    const uint32_t idesc = instr_desc_tf32(w);
                    for (int s = 0; s < d / 8; ++s) {
                        const uint64_t ad = smem_desc(a_base + c * slab_bytes + s * 1024u, d * 128u, 512u, 1u);
                        const uint64_t bd = smem_desc(e_base + c * ecb_bytes + (s >> 2) * Kpad * 128u + ch * 256u * 128u +
                                                          (s & 3) * 32u, 16u, 1024u, 2u);
                        umma_tf32(tmem_base + col, ad, bd, idesc, s > 0 ? 1u : 0u);
                    }
                    col += w;
                }
                umma_commit(bar_m);
            }
            mbar_wait(bar_m, phase_m);
            phase_m ^= 1;
            tc_fence_after();

            // ---- epilogue: one latent row per thread -----------------------------------------------------------------
            int col = 0;
            for (int u = u0; u < u1; ++u) {
                const int c = u / P.chunks, ch = u % P.chunks;
                const int w = min(256, Kpad - ch * 256);
                const int nb32 = (w + 31) >> 5;
                const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + col;
                const float* ee = ee_s + c * Kpad + ch * 256;
                const uint8_t* zrow = a_s + c * slab_bytes + warp * d * 128u;
                const uint8_t* ecb = e_s + (size_t)c * ecb_bytes;
                // |z|^2 of this row for codebook c: exact sequential chain (arithmetic contract)
                float zz = 0.0f;
                for (int j = 0; j < d; ++j) { const float v = *reinterpret_cast<const float*>(zrow + a_off(lane, j)); zz = fmaf(v, v, zz); }
                // pass 1: approximate row minimum of |e|^2 - 2 z.e
                float mn = CUDART_INF_F;
                for (int blk = 0; blk < nb32; ++blk) {
                    float v[32];
                    if (w - blk * 32 >= 32) tmem_ld32(trow + blk * 32, v); else tmem_ld16(trow + blk * 32, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int kk = blk * 32 + i;
                        if (kk < w) mn = fminf(mn, fmaf(kNeg2OverC, v[i], ee[kk]));
                        if (P.dbg && kk < w) P.dbg[((size_t)tile * kTM + tid) * (C * Kpad) + c * Kpad + kk] = v[i];
                    }
                }
                // rigorous bound on |tf32 distance - exact-chain distance|: operands truncated to 11 bits
                const float emax = emax_s[c];
                const float thr = 2.0f * (2.0f * kTf32Eps * sqrtf(zz) * 1.0001f * emax + kWinAbs * (zz + emax * emax));
                if (ch == 0) { run_mn = CUDART_INF_F; run_bv = CUDART_INF_F; run_bv2 = CUDART_INF_F; run_bi = 0x7fffffff; run_bad = false; }
                run_mn = fminf(run_mn, mn);
                const float lim = run_mn + thr;  // running minimum: a superset of the final survivor set
                // pass 2: survivors
                unsigned mask[kMaxBlk];
                int cnt = 0;
#pragma unroll
                for (int blk = 0; blk < kMaxBlk; ++blk) {
                    mask[blk] = 0u;
                    if (blk < nb32) {
                        float v[32];
                        if (w - blk * 32 >= 32) tmem_ld32(trow + blk * 32, v); else tmem_ld16(trow + blk * 32, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kk = blk * 32 + i;
                            if (kk < w && fmaf(kNeg2OverC, v[i], ee[kk]) <= lim) mask[blk] |= 1u << i;
                        }
                        cnt += __popc(mask[blk]);
                    }
                }
                if (valid) {
                    float bv;
                    int bi;
                    const bool multi = P.chunks > 1;
                    const bool last = ch == P.chunks - 1;
                    const bool finite = (zz < CUDART_INF_F) && (mn > -CUDART_INF_F) && (mn < CUDART_INF_F) && (multi || cnt >= 1);
                    if (!finite) run_bad = true;
                    if (!multi && finite && cnt == 1) {
                        bv = 0.0f; bi = 0;
#pragma unroll
                        for (int blk = 0; blk < kMaxBlk; ++blk) if (mask[blk]) bi = blk * 32 + __ffs(mask[blk]) - 1;
                    } else {
                        if (finite) {
#pragma unroll
                            for (int blk = 0; blk < kMaxBlk; ++blk) {
                                unsigned mk = mask[blk];
                                while (mk) {
                                    const int i = __ffs(mk) - 1;
                                    mk &= mk - 1;
                                    const int k = ch * 256 + blk * 32 + i;
                                    float dot = 0.0f;
                                    for (int j = 0; j < d; j += 4) {
                                        const float4 e4 = *reinterpret_cast<const float4*>(ecb + e_off(k, j, Kpad));
                                        dot = fmaf(*reinterpret_cast<const float*>(zrow + a_off(lane, j)), e4.x, dot);
                                        dot = fmaf(*reinterpret_cast<const float*>(zrow + a_off(lane, j + 1)), e4.y, dot);
                                        dot = fmaf(*reinterpret_cast<const float*>(zrow + a_off(lane, j + 2)), e4.z, dot);
                                        dot = fmaf(*reinterpret_cast<const float*>(zrow + a_off(lane, j + 3)), e4.w, dot);
                                    }
                                    const float dist = dist_f32(zz, ee_s[c * Kpad + k], dot);
                                    if (dist < run_bv) { run_bv2 = run_bv; run_bv = dist; run_bi = k; }  // ascending k: strict '<' keeps the first minimum
                                    else run_bv2 = fminf(run_bv2, dist);
                                }
                            }
                        }
                        if (last && !run_bad && run_bi != 0x7fffffff) nnear += near_tie(run_bv, run_bv2) ? 1u : 0u;
                        if (last && (run_bad || run_bi == 0x7fffffff)) {
                            // non-finite row: exact scan of every code with torch.argmin's NaN rule
                            run_bv = CUDART_INF_F; run_bi = 0x7fffffff;
                            for (int k = 0; k < K; ++k) {
                                float dot = 0.0f;
                                for (int j = 0; j < d; ++j)
                                    dot = fmaf(*reinterpret_cast<const float*>(zrow + a_off(lane, j)),
                                               *reinterpret_cast<const float*>(ecb + e_off(k, j, Kpad)), dot);
                                const float dist = dist_f32(zz, ee_s[c * Kpad + k], dot);
                                if (k == 0 || (!(dist >= run_bv) && (run_bv == run_bv))) { run_bv = dist; run_bi = k; }  // k == 0 seeds the scan (all-+inf row -> 0)
                            }
                        }
                        bv = run_bv; bi = run_bi;
                    }
                    if (last) {
                    (void)bv;
                    p.idx[seg][((size_t)b * C + c) * HW + hw] = (long long)bi;
                    // ---- fused gather + straight-through + loss --------------------------------------------------------
                    if (p.fused) {
                        float* out = p.q + ((size_t)b * C * d + (size_t)c * d) * HW + hw;
                        float ls = 0.0f;
                        for (int j = 0; j < d; j += 4) {
                            const float4 e4 = *reinterpret_cast<const float4*>(ecb + e_off(bi, j, Kpad));
                            const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float zv = *reinterpret_cast<const float*>(zrow + a_off(lane, j + t));
                                const float diff = __fsub_rn(ev[t], zv);
                                out[(size_t)(j + t) * HW] = __fadd_rn(zv, diff);
                                ls = fmaf(diff, diff, ls);
                            }
                        }
                        lsum_s[c * kThreads + tid] += ls;
                    }
                    }  // last chunk of this codebook
                }
                col += w;
            }
            tc_fence_before();
            __syncthreads();  // TMEM columns and (after the last round) the A slabs are free again
            u0 = u1;
        }
    }
    if (p.neartie) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, nnear);
        if (lane == 0 && tot) atomicAdd(p.neartie, (unsigned long long)tot);
    }
    // ---- loss: per-codebook block sums -> fp64 atomics -> last CTA finalises -------------------------------------
    if (p.fused) {
        for (int c = 0; c < C; ++c) {
            double v = (double)lsum_s[c * kThreads + tid];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp] = v;
            __syncthreads();
            if (tid == 0) atomicAdd(&p.loss_acc[c], red[0] + red[1] + red[2] + red[3]);
            __syncthreads();
        }
        __shared__ unsigned s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1u);
        __syncthreads();
        if (s_last && tid == 0) {
            __threadfence();
            float total = 0.0f;
            const double denom = (double)p.N * (double)d;
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
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

struct Plan {
    bool ok = false;
    int Kpad = 0, djb = 0, chunks = 0, units = 0, per_sm = 0;
    size_t a_bytes = 0, e_bytes = 0, smem = 0;
};

Plan make_plan(const QuantParams& p) {
    Plan pl;
    if (p.HW % 32 != 0 || p.d % 8 != 0 || p.d > 256 || p.d < 8) return pl;
    pl.Kpad = (p.K + 15) / 16 * 16;
    pl.chunks = (pl.Kpad + 255) / 256;   // codebooks wider than 256 codes take several 256-column rounds
    pl.units = p.C * pl.chunks;
    pl.djb = (p.d + 31) / 32;
    pl.a_bytes = (size_t)p.C * 4 * p.d * 128;
    pl.e_bytes = (size_t)p.C * pl.djb * pl.Kpad * 128;
    const size_t tail = sizeof(float) * ((size_t)p.C * pl.Kpad + ((p.C + 1) & ~1) + (size_t)p.C * kThreads) +
                        4 * sizeof(double) + 2 * 8 + 16;
    pl.smem = pl.a_bytes + pl.e_bytes + tail + 1024;  // + alignment slack
    if (pl.smem <= 113 * 1024) pl.per_sm = 2;
    else if (pl.smem <= 225 * 1024) pl.per_sm = 1;
    else return pl;
    for (int s = 0; s < p.n_seg; ++s)
        if (reinterpret_cast<uintptr_t>(p.z[s]) & 15) return pl;
    if (p.N > 0x7fffffffLL * 64) return pl;
    pl.ok = encode_fn() != nullptr;
    return pl;
}

float* g_dbg = nullptr;
}  // namespace

namespace tc {
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int make_maps(const QuantParams& p0, Maps& maps, int box_channels) {
    if (!encode_fn()) return CTVQ_E_UNSUPPORTED;
    for (int sg = 0; sg < p0.n_seg; ++sg) {
        const cuuint64_t dims[3] = {(cuuint64_t)p0.HW, (cuuint64_t)p0.Dtot, (cuuint64_t)p0.B};
        const cuuint64_t strides[2] = {(cuuint64_t)p0.HW * 4, (cuuint64_t)p0.HW * p0.Dtot * 4};
        const cuuint32_t box[3] = {32u, (cuuint32_t)(box_channels > 0 ? box_channels : p0.d), 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = encode_fn()(&maps.m[sg], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p0.z[sg]), dims,
                                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return CTVQ_E_UNSUPPORTED;
    }
    return CTVQ_OK;
}

int make_plain_map(CUtensorMap& m, const void* base, int dtype, long long HW, long long CH, long long B, int box_hw, int box_ch) {
    if (!encode_fn() || box_hw > 256 || box_ch > 256) return CTVQ_E_UNSUPPORTED;
    const cuuint64_t es = dtype == CTVQ_BF16 ? 2 : 4;
    const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)CH, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)HW * es, (cuuint64_t)HW * (cuuint64_t)CH * es};
    const cuuint32_t box[3] = {(cuuint32_t)box_hw, (cuuint32_t)box_ch, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = encode_fn()(&m, dtype == CTVQ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CTVQ_OK : CTVQ_E_UNSUPPORTED;
}
}  // namespace tc

extern "C" void ctvq_debug_set_tc_dump(float* buf) { g_dbg = buf; }

bool tc_supported(const QuantParams& p) { return stream_supported(p) || make_plan(p).ok; }

int launch_forward_tc(const QuantParams& p0, cudaStream_t s) {
    if (!g_dbg) {
        int rc = launch_forward_tc_fast(p0, s);
        if (rc != CTVQ_E_UNSUPPORTED) return rc;
        rc = launch_forward_tc_res(p0, s);  // single codebook resident in shared memory, 16 epilogue warps per tile
        if (rc != CTVQ_E_UNSUPPORTED) return rc;
        rc = launch_forward_tc_c1(p0, s);  // resident-codebook kernels for the configs' own small-K shapes (measured faster there)
        if (rc != CTVQ_E_UNSUPPORTED) return rc;
        rc = launch_forward_tc_stream(p0, s);  // any K: the codebook streams through a TMA ring
        if (rc != CTVQ_E_UNSUPPORTED) return rc;
    }
    const Plan pl = make_plan(p0);
    if (!pl.ok) return CTVQ_E_UNSUPPORTED;
    TcParams P;
    P.q = p0;
    P.q.tiles_per_seg = (int)((p0.N + kTM - 1) / kTM);
    P.Kpad = pl.Kpad;
    P.ntiles = P.q.tiles_per_seg * p0.n_seg;
    P.units = pl.units;
    P.chunks = pl.chunks;
    P.djb = pl.djb;
    P.a_bytes = (unsigned)pl.a_bytes;
    P.e_bytes = (unsigned)pl.e_bytes;
    P.dbg = g_dbg;
    Maps maps;
    if (make_maps(p0, maps, 0) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(vq_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return (int)e;
    int grid = sm_count() * pl.per_sm;
    if (grid > P.ntiles) grid = P.ntiles;
    vq_fwd_tc_kernel<<<grid, kThreads, pl.smem, s>>>(P, maps);
    return (int)cudaGetLastError();
}

}  // namespace ctvq
