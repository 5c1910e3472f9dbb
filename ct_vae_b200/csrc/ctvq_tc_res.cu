// Resident-codebook single-codebook tcgen05 forward kernel: VectorQuantizer of configs/vq_vae.yaml (K=512, D=64) and the
// config-4 sweep shapes with K <= 512 (D=64) / K <= 1024 (D=32).  Replaces models/vq_vae.py:25-55 and
// models/mcq_vae.py:26-74 for C = 1.
//
// Why a second single-codebook kernel: the streaming kernel (ctvq_tc_stream.cu) re-reads the codebook from L2 for every
// super-tile, gathers code rows from L2 for exact re-scoring and gives each latent row to ONE thread for all K codes, so an
// SM runs 8 epilogue warps at 168 registers (round-1 ncu: issue active 37 %, 0.21 of either roof at config 1).  Here
//   * the whole codebook sits in shared memory ONCE per persistent CTA, TMA-loaded straight into the K-major
//     SWIZZLE_128B UMMA layout (2-D tensor map over the nn.Parameter; rows >= K zero-filled by the TMA unit); the same
//     copy serves the tensor core, the exact re-scoring and the final gather;
//   * a tile is 128 latent rows x all K codes, scored as NH units of 256 codes into a DOUBLE-BUFFERED accumulator
//     (2 x 256 = all 512 TMEM columns): the MMAs of unit u+1 run while the epilogue filters unit u;
//   * ALL 16 epilogue warps work on the SAME tile: warp (q, s) owns the 32 rows of TMEM lane quarter q and the 64-column
//     slice s of every unit; the four slices of a row meet through 128-thread named barriers (one per unit for the slice
//     maxima, two for the exact scores).  Each thread keeps only the D/4 channels of its row that it gathers/stores at
//     the end (16 registers at D=64), so 544 threads fit in the register file without spills;
//   * |e_k|^2 rides in the GEMM as one extra K-group (A = constant ones, B = -|e_k|^2/2 as three tf32 terms, an MN-major
//     operand: 8 KB per unit), so the accumulator IS the score z.e_k - |e_k|^2/2 and the per-(row, code) work is
//         pass 1   3-input max                                         (FMNMX3: 0.5 instructions)
//         pass 2   survivor bitmask  s_k >= running max - bound/2      (FSETP + predicated LOP: 2 instructions)
//   * every slice sees all (unit, slice) maxima, so it knows without a further exchange whether its row has ONE candidate
//     slice; a lone candidate wins unscored.  All other candidates go to a per-quarter LIST in shared memory and are
//     scored exactly ONE (row, code) PAIR PER LANE -- the SIMT-divergent "each thread re-scores its own survivors" loop of
//     the other kernels costs a full warp pass for one or two active lanes (first version of this kernel: 268 M warp
//     instructions, half of them in that loop).  The per-row winner is a 64-bit atomicMin over (distance, code) keys whose
//     unsigned order is torch.argmin's (NaN first, ties -> lower index); near-tie flags compare each scored pair with it.
// The arithmetic contract is the same as every other kernel's: indices equal the C oracle's on every row.
#include "ctvq_tc_ptx.cuh"

namespace ctvq {
using namespace tc;
namespace {

constexpr int kEW = 16;                  // epilogue warps
constexpr int kRT = 32 * kEW + 64;       // + the MMA-issue warp and the TMA-issue warp (one lane each)

struct ResParams {
    QuantParams q;
    int ntiles;
    unsigned hw_mul, hw_shift;   // n / HW for 32-bit n as (((n - t) >> 1) + t) >> hw_shift, t = umulhi(n, hw_mul)
    const float* ee;             // [NK] exact |e_k|^2 (+inf for padded codes), written by res_prep_kernel
    const unsigned* emax_bits;   // max |e_k|^2 as float bits (NaN / inf poison the bound)
};

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tma_load_2d_r(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void or_if_ge_r(unsigned& m, float a, float lim, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(m) : "f"(a), "f"(lim), "r"(bit));
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {  // global -> L2 only
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ unsigned fast_div(unsigned n, unsigned mul, unsigned shift) {
    const unsigned t = __umulhi(n, mul);
    return (((n - t) >> 1) + t) >> shift;
}
__device__ __forceinline__ float sqrt_approx_r(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Tiles requested into L2 ahead of their shared-memory load.  0 = the L2 request of a tile goes out right before the TMA
// lane starts waiting for the tile's ring slot, i.e. the lead is exactly the time the slot is still busy.  Re-measured after
// the window was centred (less epilogue time per tile): 2 tiles ahead -- the round-2 setting -- had become a liability at
// config 3 (K=64, d=128: 0.256 ms back to back / 0.209 ms behind a backward, against 0.199 ms in every context with 0;
// 3 and 4 tiles ahead: 0.27-0.29 ms): lines requested too early are evicted by the kernel's own 0.5 GB write stream
// before the load arrives, and then cross HBM twice.  No request at all is 3 % slower than 0 at H*W = 256.
constexpr int kL2Ahead = 0;
constexpr int kCap = 128;  // pairs per (tile parity, lane quarter) list: one per thread of the four slice warps

// (distance, code) as ONE 64-bit key whose unsigned order is torch.argmin's: NaN first, then ascending distance
// (-inf ... +inf), ties -> lower code index
__device__ __forceinline__ unsigned long long pack_key(float d, int k) {
    const unsigned u = __float_as_uint(d);
    const unsigned o = (d != d) ? 0u : ((u & 0x80000000u) ? ~u : (u | 0x80000000u));
    return ((unsigned long long)o << 32) | (unsigned)k;
}
__device__ __forceinline__ float unpack_dist(unsigned long long key) {
    const unsigned o = (unsigned)(key >> 32);
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// Pre-pass, one thread per (padded) code: exact |e_k|^2 (sequential chain of the arithmetic contract) and its maximum.
__global__ void res_prep_kernel(const float* __restrict__ E, int K, int D, int NK, float* __restrict__ ee,
                                unsigned* __restrict__ emax_bits) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= NK) return;
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
}

// D: channels (32 / 64 / 128); NKU: codes per unit = accumulator columns per buffer (256, or 64 for small codebooks);
// NH: units per tile (K <= NKU*NH); NSTAGE: slab ring depth
template <int D, int NKU, int NH, int NSTAGE>
__global__ void __launch_bounds__(kRT, 1) vq_fwd_tc_res_kernel(const ResParams P, const __grid_constant__ Maps maps,
                                                               const __grid_constant__ CUtensorMap emap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const QuantParams& p = P.q;
    const int K = p.K, HW = p.HW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NK = NH * NKU;
    constexpr int SL = NKU / 4;                      // columns of one warp's slice of a unit (64 or 16)
    constexpr int KB = D / 32;                       // 128-byte K-blocks of the codebook operand
    constexpr int CH = D / 4;                        // channels of a thread's gather slice
    constexpr uint32_t kBlk = (uint32_t)D * 128u;    // one 32-row block of a slab: [D][128 B]
    constexpr uint32_t kStage = 4u * kBlk;
    constexpr uint32_t kBbytes = (uint32_t)KB * NK * 128u;
    constexpr uint32_t kXunit = (uint32_t)(NKU / 32) * 1024u;  // extra-K-group operand of one unit: [NKU/32 groups of 32 codes][8 k][128 B]
    static_assert(D == 32 || D == 64 || D == 128, "channel slices of 8 / 16 / 32 per warp");
    static_assert(NKU == 256 || NKU == 64, "slices of 64 or 16 columns");
    static_assert(NSTAGE >= 2, "ring");
    uint8_t* a_s = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_s = a_s + (size_t)NSTAGE * kStage;
    uint8_t* x_s = b_s + kBbytes;
    uint8_t* ones_s = x_s + (size_t)NH * kXunit;     // [4 row blocks][8 k][32 rows]: A operand of the extra K-group
    float* pm_s = reinterpret_cast<float*>(ones_s + 4096);   // [2][4][128] slice maxima of the unit (parity-buffered)
    float* pz_s = pm_s + 2 * 4 * 128;                         // [4][128] partial |z|^2 per slice
    unsigned long long* key_s = reinterpret_cast<unsigned long long*>(pz_s + 4 * 128);  // [2][128] (distance, code) minimum per row
    unsigned* near_s = reinterpret_cast<unsigned*>(key_s + 2 * 128);                    // [2][128] near-tie flags
    unsigned* list_s = near_s + 2 * 128;                      // [2][4][kCap] (row << 16 | code) pairs awaiting an exact score
    unsigned* listn_s = list_s + 2 * 4 * kCap;                // [2][4] (+ pad to 8 bytes)
    uint64_t* bars = reinterpret_cast<uint64_t*>(listn_s + 8);
    // full[NSTAGE] empty[NSTAGE] mma[2] tfree[2] bfull
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 5);
    unsigned* s_last = tmem_slot + 1;
    const uint32_t a_base = smem_u32(a_s), b_base = smem_u32(b_s), x_base = smem_u32(x_s), ones_base = smem_u32(ones_s);
    const uint32_t bar_full0 = smem_u32(&bars[0]), bar_empty0 = bar_full0 + 8 * NSTAGE;
    const uint32_t bar_m = bar_empty0 + 8 * NSTAGE, bar_tfree = bar_m + 16, bar_bfull = bar_tfree + 16;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full0 + 8 * i, 1); mbar_init(bar_empty0 + 8 * i, kEW + 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_m + 8 * i, 1); mbar_init(bar_tfree + 8 * i, kEW); }
        mbar_init(bar_bfull, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    __syncthreads();  // barriers initialised before the first TMA may signal them

    const int niter = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool producer = (tid == 32 * (kEW + 1));  // the TMA-issue lane
    const unsigned tps = (unsigned)p.tiles_per_seg;
    auto tile_pos = [&](int it, int& seg, unsigned& row0) {  // tile of iteration `it` -> (segment, first row); N < 2^31 rows per segment
        const unsigned tile = blockIdx.x + (unsigned)it * gridDim.x;
        seg = (p.n_seg == 1) ? 0 : (int)(tile / tps);
        row0 = (tile - (unsigned)seg * tps) * (unsigned)kTM;
    };
    auto issue_tma = [&](int it, bool l2_only) {  // TMA lane: load the tile of iteration `it` into its ring slot / into L2
        int seg;
        unsigned row0;
        tile_pos(it, seg, row0);
        const int st = it % NSTAGE;
        int nblk = 0;
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) nblk += ((long long)row0 + 32 * mb < p.N) ? 1 : 0;
        if (!l2_only) mbar_expect_tx(bar_full0 + 8 * st, (uint32_t)nblk * kBlk);
        for (int mb = 0; mb < nblk; ++mb) {
            const unsigned nb = row0 + 32u * mb;
            const unsigned bb = fast_div(nb, P.hw_mul, P.hw_shift);
            if (l2_only) tma_prefetch_3d(&maps.m[seg], (int)(nb - bb * (unsigned)HW), 0, (int)bb);
            else tma_load_3d(a_base + st * kStage + mb * kBlk, &maps.m[seg], bar_full0 + 8 * st, (int)(nb - bb * (unsigned)HW), 0, (int)bb);
        }
    };
    if (producer) {
        // the codebook: KB x NK/64 boxes of [64 codes x 32 channels] land as the K-major SWIZZLE_128B operand
        mbar_expect_tx(bar_bfull, kBbytes);
        for (int kb = 0; kb < KB; ++kb)
            for (int i = 0; i < NK / 64; ++i)
                tma_load_2d_r(b_base + (uint32_t)kb * NK * 128u + (uint32_t)i * 8192u, &emap, bar_bfull, kb * 32, i * 64);
        for (int it = 0; it < NSTAGE && it < niter; ++it) issue_tma(it, false);
        for (int it = NSTAGE; it < NSTAGE + kL2Ahead && it < niter; ++it) issue_tma(it, true);
    }
    // extra K-group operands: x[unit][code][k] = -|e|^2/2 as three tf32-exact terms (k = 0..2), zero for k = 3..7
    for (int n = tid; n < NK; n += kRT) {
        const float a = __ldg(P.ee + n);
        float t[8] = {-1.0e30f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};  // padded / overflowed codes never survive the filter
        if (a < CUDART_INF_F) {
            const float h = (-0.5f * kTruncC) * a;  // centred truncation error (ctvq_common.cuh)
            t[0] = __uint_as_float(__float_as_uint(h) & 0xFFFFE000u);
            const float r1 = h - t[0];
            t[1] = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
            t[2] = __uint_as_float(__float_as_uint(r1 - t[1]) & 0xFFFFE000u);
        }
        const int m = n % NKU;
        uint8_t* blk = x_s + (size_t)(n / NKU) * kXunit + (size_t)(m >> 5) * 1024;
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float*>(blk + a_off(m & 31, j)) = t[j];
    }
    for (int i = tid; i < 1024; i += kRT)  // rows are constant, so the swizzle inside a 128-byte row is immaterial
        reinterpret_cast<float*>(ones_s)[i] = ((i >> 5) & 7) < 3 ? 1.0f : 0.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const float emax = sqrtf(__uint_as_float(__ldg(P.emax_bits))) * 1.0001f;

    float lsum = 0.0f;
    unsigned nnear = 0u;  // near-tie rows seen by this thread (include/ctvq.h)
    if (warp == kEW + 1) {
        // =============================== TMA lane: slab ring + L2 requests (kL2Ahead, see its comment) ===============
        // A slab can only be re-filled once every epilogue warp has scored its last candidate from it, which leaves less
        // than a tile time of lookahead with the 2 stages that fit beside a 128 KB codebook -- not enough to cover HBM
        // latency.  So every tile is ALSO requested into L2 (cp.async.bulk.prefetch.tensor) BEFORE the lane starts waiting
        // for its ring slot: the shared-memory load that follows the wait then completes at L2 latency.  HBM traffic stays
        // one read per slab as long as the request is not issued too early (kL2Ahead above).
        if (lane == 0) {
            for (int nx = NSTAGE; nx < niter; ++nx) {
                const int prev = nx - NSTAGE;
                if (nx + kL2Ahead < niter) issue_tma(nx + kL2Ahead, true);
                mbar_wait_sleep(bar_empty0 + 8 * (prev % NSTAGE), (uint32_t)(prev / NSTAGE) & 1u);
                issue_tma(nx, false);
            }
        }
    } else if (warp == kEW) {
        // =============================== MMA lane ====================================================================
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32(NKU);
            const uint32_t idesc_x = idesc | (1u << 16);  // extra K-group: B operand MN-major as well
            mbar_wait_sleep(bar_bfull, 0u);
            for (int it = 0; it < niter; ++it) {
                const int st = it % NSTAGE;
                const uint32_t stage_u32 = a_base + st * kStage;
                mbar_wait_sleep(bar_full0 + 8 * st, (uint32_t)(it / NSTAGE) & 1u);
#pragma unroll 1
                for (int h = 0; h < NH; ++h) {
                    const int u = it * NH + h, buf = u & 1;
                    if (u >= 2) mbar_wait_sleep(bar_tfree + 8 * buf, (uint32_t)((u >> 1) - 1) & 1u);
                    tc_fence_after();
                    const uint32_t dcol = tmem_base + buf * NKU;
#pragma unroll
                    for (int s = 0; s < D / 8; ++s) {
                        const uint64_t ad = smem_desc(stage_u32 + (uint32_t)s * 1024u, kBlk, 512u, 1u);
                        const uint64_t bd = smem_desc(b_base + (uint32_t)(s >> 2) * NK * 128u + (uint32_t)h * NKU * 128u + (s & 3) * 32u,
                                                      16u, 1024u, 2u);
                        umma_tf32(dcol, ad, bd, idesc, s > 0 ? 1u : 0u);
                    }
                    // + 1 * (-|e_k|^2 / 2): the accumulator now holds the whole score z.e_k - |e_k|^2/2
                    umma_tf32(dcol, smem_desc(ones_base, 1024u, 512u, 1u), smem_desc(x_base + (uint32_t)h * kXunit, 1024u, 512u, 1u),
                              idesc_x, 1u);
                    umma_commit(bar_m + 8 * buf);
                }
                umma_commit(bar_empty0 + 8 * st);  // every MMA that reads this slab has been issued before this commit
            }
        }
    } else {
        // =============================== epilogue warp (q, s): rows of lane quarter q, column slice s ===============
        const int q = warp & 3, s = warp >> 2;
        const int r = q * 32 + lane;  // row within the tile
        uint32_t zsw[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) zsw[x] = ((((lane >> 3) ^ x) & 3) << 5) + ((lane & 7) << 2);
        mbar_wait_fast(bar_bfull, 0u);  // the codebook copy is read through the generic proxy below
        // exact fp32 distance of (row rl of this quarter's slab block, code k): arithmetic contract of DESIGN.md
        auto exact_dist = [&](const uint8_t* zblk, int rl, int k, float& zz_out) -> float {
            const uint32_t sw0 = (uint32_t)((rl & 7) << 2), rh = (uint32_t)(rl >> 3);
            const uint8_t* erow = b_s + k * 128;
            const uint32_t kx = (uint32_t)(k & 7) << 4;
            float dot = 0.0f, zz = 0.0f;
#pragma unroll
            for (int j = 0; j < D; j += 4) {
                const float4 e4 = *reinterpret_cast<const float4*>(erow + (j >> 5) * NK * 128 + ((((j & 31) >> 2) << 4) ^ kx));
                const float z0 = *reinterpret_cast<const float*>(zblk + j * 128 + ((((rh ^ j) & 3)) << 5) + sw0);
                const float z1 = *reinterpret_cast<const float*>(zblk + (j + 1) * 128 + ((((rh ^ (j + 1)) & 3)) << 5) + sw0);
                const float z2 = *reinterpret_cast<const float*>(zblk + (j + 2) * 128 + ((((rh ^ (j + 2)) & 3)) << 5) + sw0);
                const float z3 = *reinterpret_cast<const float*>(zblk + (j + 3) * 128 + ((((rh ^ (j + 3)) & 3)) << 5) + sw0);
                zz = fmaf(z0, z0, zz); dot = fmaf(z0, e4.x, dot);
                zz = fmaf(z1, z1, zz); dot = fmaf(z1, e4.y, dot);
                zz = fmaf(z2, z2, zz); dot = fmaf(z2, e4.z, dot);
                zz = fmaf(z3, z3, zz); dot = fmaf(z3, e4.w, dot);
            }
            zz_out = zz;
            return dist_f32(zz, __ldg(P.ee + k), dot);
        };
        for (int it = 0; it < niter; ++it) {
            int seg;
            unsigned row0;
            tile_pos(it, seg, row0);
            const unsigned n = row0 + (unsigned)r;
            const bool valid = (long long)n < p.N;  // uniform over the four warps of a quarter (N is a multiple of 32)
            const unsigned b = fast_div(n, P.hw_mul, P.hw_shift);
            const int hw = (int)(n - b * (unsigned)HW);
            const int st = it % NSTAGE, par = it & 1;
            unsigned long long* keyp = key_s + par * 128;
            unsigned* nearp = near_s + par * 128;
            unsigned* listp = list_s + (par * 4 + q) * kCap;
            unsigned* listn = listn_s + par * 4 + q;
            mbar_wait_fast(bar_full0 + 8 * st, (uint32_t)(it / NSTAGE) & 1u);
            const uint8_t* zblk = a_s + st * kStage + q * kBlk;
            float zreg[CH];  // the channels this thread gathers / stores at the end
            if (valid) {
                float pz = 0.0f;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    zreg[i] = *reinterpret_cast<const float*>(zblk + (CH * s + i) * 128 + zsw[(CH * s + i) & 3]);
                    pz = fmaf(zreg[i], zreg[i], pz);
                }
                pz_s[s * 128 + r] = pz;  // partial |z|^2: only an UPPER BOUND of |z|^2 is needed by the filter
            }
            if (s == 0) { keyp[r] = ~0ull; nearp[r] = 0u; }
            if (s == 0 && lane == 0) *listn = 0u;

            float run = -CUDART_INF_F, top2 = -CUDART_INF_F, thr = 0.0f, zzu = 0.0f;
            float pm[NH];
            unsigned mlo[NH], mhi[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const int u = it * NH + h, buf = u & 1;
                mbar_wait_fast(bar_m + 8 * buf, (uint32_t)(u >> 1) & 1u);
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * NKU + s * SL;
                constexpr int LW = SL >= 32 ? 32 : SL;   // columns per TMEM load
                uint32_t a[LW];
                float pmh = -CUDART_INF_F;
                if (valid) {
                    // ---- pass 1: maximum of this slice's SL approximate scores --------------------------------------
                    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
#pragma unroll
                    for (int hh = 0; hh < SL / LW; ++hh) {
                        tmem_ld_issue(trow + LW * hh, a);
                        tmem_ld_wait(a);
#pragma unroll
                        for (int i = 0; i < LW; i += 4) {
                            m0 = fmaxf(m0, __uint_as_float(a[i])); m1 = fmaxf(m1, __uint_as_float(a[i + 1]));
                            m2 = fmaxf(m2, __uint_as_float(a[i + 2])); m3 = fmaxf(m3, __uint_as_float(a[i + 3]));
                        }
                    }
                    pmh = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    pm_s[((u & 1) * 4 + s) * 128 + r] = pmh;
                }
                named_sync(1 + q, 128);  // the four slices of these 32 rows
                mlo[h] = 0u; mhi[h] = 0u;
                if (valid) {
                    const float* pmr = pm_s + (u & 1) * 4 * 128 + r;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {  // running maximum and runner-up over all (unit, slice) maxima
                        const float v = pmr[w * 128];
                        top2 = fmaxf(top2, fminf(run, v));
                        run = fmaxf(run, v);
                    }
                    if (h == 0) {
                        zzu = ((pz_s[r] + pz_s[128 + r]) + (pz_s[256 + r] + pz_s[384 + r])) * 1.00001f;  // >= the exact chain value
                        // rigorous bound on |tf32 distance - exact-chain distance|; scores are distances / -2
                        thr = 2.0f * (2.0f * kTf32Eps * sqrt_approx_r(zzu) * 1.0001f * emax + kWinAbs * (zzu + emax * emax));
                    }
                    const float lim = run - 0.5f * thr;  // running maximum: a superset of the final survivor set
                    // ---- pass 2: survivors as a bitmask (both halves re-read from TMEM: nothing lives across the barrier) --
#pragma unroll
                    for (int hh = 0; hh < SL / LW; ++hh) {
                        tmem_ld_issue(trow + LW * hh, a);
                        tmem_ld_wait(a);
                        unsigned mk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                        for (int i = 0; i < LW; ++i) or_if_ge_r(mk[i & 3], __uint_as_float(a[i]), lim, 1u << i);
                        const unsigned m = (mk[0] | mk[1]) | (mk[2] | mk[3]);
                        if (hh == 0) mlo[h] = m; else mhi[h] = m;
                    }
                    // (an fma-pipe form of this pass -- t = sat((s - lim') * 2^80), bits summed in fp32 accumulators, two
                    // immediate-form FFMAs per score -- was measured SLOWER: 0.275 vs 0.248 ms at K=512, 1 M rows; the kernel
                    // is bound by its barrier / latency structure at 4 warps per scheduler, not by the alu pipe)
                }
                pm[h] = pmh;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tfree + 8 * buf);  // accumulator drained: the MMAs of unit u+2 may start
            }
            // ---- decide: every slice knows all (unit, slice) maxima, hence whether the row has ONE candidate slice ----------
            const float lim = run - 0.5f * thr;
            int cnt = 0;
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                if (pm[h] < lim) { mlo[h] = 0u; mhi[h] = 0u; }  // the whole slice fell out of the FINAL window
                cnt += __popc(mlo[h]) + __popc(mhi[h]);
            }
            // non-finite rows (NaN / inf latents, overflowing norms, a poisoned bound): same verdict in all four slices
            const bool bad = !(zzu < CUDART_INF_F) || !(run > -CUDART_INF_F) || !(run < CUDART_INF_F) || !(lim == lim);
            const bool multi = top2 >= lim;  // a second (unit, slice) maximum inside the window: the row needs exact scores
            float fd1 = CUDART_INF_F, fd2 = CUDART_INF_F;  // fallback path: exact best / second-best of this thread's candidates
            int fk1 = -1;
            bool fallback = false;
            if (valid) {
                if (bad) {
                    // exact scan of this slice's codes with torch.argmin's NaN rule (first NaN wins; an all-+inf row answers
                    // the lowest index): the key's order (NaN < -inf < ... < +inf, then index) does the rest
#pragma unroll 1
                    for (int h = 0; h < NH; ++h)
#pragma unroll 1
                        for (int i = 0; i < SL; ++i) {
                            const int k = h * NKU + s * SL + i;
                            if (k >= K) break;
                            float zz;
                            const float dist = exact_dist(zblk, lane, k, zz);
                            atomicMin(keyp + r, pack_key(dist, k));
                        }
                } else if (cnt >= 1) {
                    if (!multi && cnt == 1) {  // the only code in the final window: no exact distance needed
                        int k = 0;
#pragma unroll
                        for (int h = 0; h < NH; ++h) {
                            if (mlo[h]) k = h * NKU + s * SL + __ffs(mlo[h]) - 1;
                            if (mhi[h]) k = h * NKU + s * SL + 32 + __ffs(mhi[h]) - 1;
                        }
                        keyp[r] = (unsigned long long)(unsigned)k;
                    } else {
                        // hand the candidates to the quarter's shared list: they are scored one PAIR PER LANE, all lanes busy
                        const unsigned base = atomicAdd(listn, (unsigned)cnt);
                        if (base + (unsigned)cnt <= (unsigned)kCap) {
                            unsigned e = base;
#pragma unroll
                            for (int h = 0; h < NH; ++h) {
                                unsigned long long mk = ((unsigned long long)mhi[h] << 32) | mlo[h];
                                while (mk) {
                                    const int i = __ffsll((long long)mk) - 1;
                                    mk &= mk - 1;
                                    listp[e++] = ((unsigned)lane << 16) | (unsigned)(h * NKU + s * SL + i);
                                }
                            }
                        } else {
                            // list full (pathologically tie-heavy tile): this thread scores its own candidates
                            for (unsigned e = base; e < (unsigned)kCap; ++e) listp[e] = 0xFFFFFFFFu;
                            fallback = true;
#pragma unroll 1
                            for (int h = 0; h < NH; ++h) {
                                unsigned long long mk = ((unsigned long long)mhi[h] << 32) | mlo[h];
                                while (mk) {
                                    const int k = h * NKU + s * SL + __ffsll((long long)mk) - 1;
                                    mk &= mk - 1;
                                    float zz;
                                    const float dist = exact_dist(zblk, lane, k, zz);
                                    atomicMin(keyp + r, pack_key(dist, k));
                                    if (dist < fd1) { fd2 = fd1; fd1 = dist; fk1 = k; } else fd2 = fminf(fd2, dist);
                                }
                            }
                        }
                    }
                }
            }
            named_sync(1 + q, 128);  // the list is complete
            // ---- phase 1: one (row, code) pair per lane, exact fp32 distance, lexicographic (distance, index) minimum per row ---
            int pr = -1, pk = 0;
            float pd = 0.0f;
            {
                const unsigned nl = min(*listn, (unsigned)kCap);
                const unsigned e = (unsigned)(s * 32 + lane);
                if (e < nl) {
                    const unsigned ent = listp[e];
                    if (ent != 0xFFFFFFFFu) {
                        pr = (int)(ent >> 16); pk = (int)(ent & 0xFFFFu);
                        float zz;
                        pd = exact_dist(zblk, pr, pk, zz);
                        atomicMin(keyp + q * 32 + pr, pack_key(pd, pk));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty0 + 8 * st);  // this warp no longer needs the slab
            named_sync(1 + q, 128);  // every candidate of these 32 rows has been scored
            // ---- phase 2: near-tie flags (relative gap to the winner, include/ctvq.h), then the winner itself ---------------
            if (pr >= 0) {
                const unsigned long long w = keyp[q * 32 + pr];
                if ((int)(unsigned)w != pk && near_tie(unpack_dist(w), pd)) nearp[q * 32 + pr] = 1u;
            }
            if (fallback) {
                const unsigned long long w = keyp[r];
                const float cand = ((int)(unsigned)w == fk1) ? fd2 : fd1;
                if (near_tie(unpack_dist(w), cand)) nearp[r] = 1u;
            }
            if (valid) {
                const int bi = (int)(unsigned)keyp[r];
                if (s == 0) p.idx[seg][(size_t)b * HW + hw] = (long long)bi;
                // ---- fused gather + straight-through + loss for this thread's CH channels -------------------------------------
                if (p.fused) {
                    float* out = p.q + ((size_t)b * D + CH * s) * HW + hw;  // a warp's 32 rows are contiguous: 128-byte stores
                    const uint8_t* erow = b_s + bi * 128 + ((CH * s) >> 5) * NK * 128;
                    const uint32_t kx = (uint32_t)(bi & 7) << 4;
                    float ls0 = 0.0f, ls1 = 0.0f;
#pragma unroll
                    for (int j = 0; j < CH; j += 4) {
                        const float4 e4 = *reinterpret_cast<const float4*>(erow + (((((CH * s + j) & 31) >> 2) << 4) ^ kx));
                        const float d0 = __fsub_rn(e4.x, zreg[j]), d1 = __fsub_rn(e4.y, zreg[j + 1]);
                        const float d2 = __fsub_rn(e4.z, zreg[j + 2]), d3 = __fsub_rn(e4.w, zreg[j + 3]);
                        out[(size_t)j * HW] = __fadd_rn(zreg[j], d0);  // z + (q - z), models/vq_vae.py:53
                        out[(size_t)(j + 1) * HW] = __fadd_rn(zreg[j + 1], d1);
                        out[(size_t)(j + 2) * HW] = __fadd_rn(zreg[j + 2], d2);
                        out[(size_t)(j + 3) * HW] = __fadd_rn(zreg[j + 3], d3);
                        ls0 = fmaf(d0, d0, ls0); ls1 = fmaf(d1, d1, ls1);
                        ls0 = fmaf(d2, d2, ls0); ls1 = fmaf(d3, d3, ls1);
                    }
                    lsum += ls0 + ls1;
                }
            }
            // near-tie flags of the PREVIOUS tile are complete now (their writers passed two barriers since): count them
            if (s == 0 && it > 0) nnear += near_s[(par ^ 1) * 128 + r];
        }
        // the last tile's flags: one more barrier orders the phase-2 writes before the count
        named_sync(1 + q, 128);
        if (s == 0 && niter > 0) nnear += near_s[((niter - 1) & 1) * 128 + r];
    }
    if (p.neartie && warp < kEW) {
        const unsigned tot = __reduce_add_sync(0xffffffffu, nnear);
        if (lane == 0 && tot) atomicAdd(p.neartie, (unsigned long long)tot);
    }
    // ---- loss: warp sums -> fp64 atomics -> last CTA finalises -----------------------------------------------------
    if (p.fused) {
        if (warp < kEW) {
            double v = (double)lsum;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v != 0.0) atomicAdd(&p.loss_acc[0], v);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) *s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1u);
        __syncthreads();
        if (*s_last && tid == 0) {
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

int make_map_codebook(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows) {
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

template <int D, int NKU, int NH, int NSTAGE>
constexpr size_t res_smem() {
    return (size_t)NSTAGE * 4 * D * 128 + (size_t)(D / 32) * NH * NKU * 128 + (size_t)NH * (NKU / 32) * 1024 + 4096 +
           sizeof(float) * (2 * 4 * 128 + 4 * 128) + 2 * 128 * 8 + 2 * 128 * 4 + (2 * 4 * kCap + 8) * 4 + (2 * NSTAGE + 5) * 8 + 16 + 1024;
}

template <int D, int NKU, int NH, int NSTAGE>
int launch_res(const QuantParams& p0, cudaStream_t s) {
    constexpr int NK = NH * NKU;
    // scratch: [emax bits, pad to 256 B][ee: NK floats]
    const size_t need = 256 + (size_t)NK * 4;
    if (!p0.scratch || p0.scratch_bytes < need || (reinterpret_cast<uintptr_t>(p0.scratch) & 255)) return CTVQ_E_UNSUPPORTED;
    unsigned* emax_bits = reinterpret_cast<unsigned*>(p0.scratch);
    float* ee = reinterpret_cast<float*>(p0.scratch + 256);
    ResParams P;
    P.q = p0;
    P.q.tiles_per_seg = (int)((p0.N + kTM - 1) / kTM);
    P.ntiles = P.q.tiles_per_seg * p0.n_seg;
    P.ee = ee;
    P.emax_bits = emax_bits;
    {   // n / HW by multiply-shift (HW >= 32 here): l = ceil(log2 HW), mul = floor(2^32 (2^l - HW) / HW) + 1
        unsigned l = 0;
        while ((1u << l) < (unsigned)p0.HW) ++l;
        P.hw_mul = (unsigned)((((unsigned long long)1 << 32) * ((1ull << l) - (unsigned)p0.HW)) / (unsigned)p0.HW + 1);
        P.hw_shift = l - 1;
    }
    Maps maps;
    if (make_maps(p0, maps, D) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    CUtensorMap emap;
    if (make_map_codebook(&emap, p0.E[0], (uint64_t)D, (uint64_t)p0.K) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    cudaError_t e = cudaMemsetAsync(emax_bits, 0, 4, s);
    if (e != cudaSuccess) return (int)e;
    res_prep_kernel<<<(NK + 255) / 256, 256, 0, s>>>(p0.E[0], p0.K, D, NK, ee, emax_bits);
    constexpr size_t smem = res_smem<D, NKU, NH, NSTAGE>();
    static_assert(smem <= 227 * 1024, "one CTA per SM");
    auto kern = vq_fwd_tc_res_kernel<D, NKU, NH, NSTAGE>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int grid = sm_count();
    if (grid > P.ntiles) grid = P.ntiles;
    kern<<<grid, kRT, smem, s>>>(P, maps, emap);
    return (int)cudaGetLastError();
}

// (NKU, NH) for a shape, 0 when the kernel does not cover it
static void res_plan(int d, int K, int& nku, int& nh) {
    nku = nh = 0;
    if (K <= 64 && (d == 64 || d == 128)) { nku = 64; nh = 1; return; }
    const int units = (K + 255) / 256;
    if (d == 64 && units <= 2) { nku = 256; nh = units; }
    else if (d == 32 && units <= 4) { nku = 256; nh = units == 3 ? 4 : units; }
}

}  // namespace

bool res_supported(const QuantParams& p) {
    if (p.C != 1 || p.HW % 32 != 0 || p.d != p.Dtot) return false;
    int nku, nh;
    res_plan(p.d, p.K, nku, nh);
    if (!nku) return false;
    if (p.N >= (1ll << 31) - 256 || p.HW > (1 << 20)) return false;  // 32-bit row arithmetic inside the kernel
    for (int sg = 0; sg < p.n_seg; ++sg)
        if (reinterpret_cast<uintptr_t>(p.z[sg]) & 15) return false;
    if (reinterpret_cast<uintptr_t>(p.E[0]) & 15) return false;
    const size_t need = 256 + (size_t)nku * nh * 4;
    if (!p.scratch || p.scratch_bytes < need || (reinterpret_cast<uintptr_t>(p.scratch) & 255)) return false;
    return encode_fn() != nullptr;
}

int launch_forward_tc_res(const QuantParams& p, cudaStream_t s) {
    if (!res_supported(p)) return CTVQ_E_UNSUPPORTED;
    int nku, nh;
    res_plan(p.d, p.K, nku, nh);
    if (nku == 64) return p.d == 128 ? launch_res<128, 64, 1, 2>(p, s) : launch_res<64, 64, 1, 4>(p, s);
    if (p.d == 64) return nh == 1 ? launch_res<64, 256, 1, 3>(p, s) : launch_res<64, 256, 2, 2>(p, s);
    if (nh == 1) return launch_res<32, 256, 1, 4>(p, s);
    if (nh == 2) return launch_res<32, 256, 2, 4>(p, s);
    return launch_res<32, 256, 4, 3>(p, s);
}

}  // namespace ctvq
