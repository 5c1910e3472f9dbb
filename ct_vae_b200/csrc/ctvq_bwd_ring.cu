// Single-codebook backward for codebooks of MANY codes whose [K, D] accumulator still fits in shared memory:
// VectorQuantizer of configs/vq_vae.yaml (K=512, D=64) and the config-4 sweep shapes with K*D <= 32768.
// Replaces autograd through models/vq_vae.py:43-53 (one_hot^T @ (q - z), the straight-through and commitment terms).
//
// vq_bwd_c1_kernel (ctvq_bwd_c1.cu) runs this shape as "stage a tile, barrier, accumulate, barrier, grad_z" with
// global fp32 atomics per (row, channel) and reaches 30 % of the HBM roofline (round-2 ncu: issue active 39 %, every
// phase waits for the slowest load of the previous one): 0.328 ms at config 1, 1 M rows.  Here one persistent CTA per
// SM keeps the WHOLE [K, D] accumulator in shared memory (128 KB at config 1) and nothing on the tile loop's critical
// path waits for HBM (0.192 ms, 65 %):
//   * 8 "worker" warps stage z with 8-byte cp.async copies into a padded [D][66] tile, THREE buffers, two tiles ahead
//     (completion on an mbarrier) -- and every tile is requested into L2 five tiles ahead by ONE tensor-map prefetch, so the
//     staging lead only has to cover L2 latency; the tile's indices ride one tile ahead in a register, range-checked once;
//   * the same warps turn the tile into q - z IN PLACE: lanes along the channel, 8 rows x D/32 chunks of coalesced
//     128-byte codebook reads from L2 per lane (the codebook does not fit beside the accumulator), issued one tile
//     ahead so they fly under the previous tile's grad_z pass;
//   * 8 "acc" warps add q - z into the accumulator with plain read-modify-write: each owns (32-channel chunk, code
//     residue class k mod R), so no two warps ever touch the same word -- race-free without atomics (fp32 shared
//     atomics are a compare-and-swap loop on sm_100a).  A warp compacts its class's rows into a list (ballot + prefix
//     popcount; a find-first-set chain measured 3x slower), then walks it four rows at a time, software-pipelined two
//     groups deep, loads before stores when the four codes differ;
//   * the worker warps then compute grad_z = g_out - coef_z (q - z) from shared memory only: g_out streams through a
//     TMA ring, ONE 3-D tensor-map box [D][64] per tile (a 256-byte bulk copy per channel costs ~46 cycles of TMA
//     service each: measured 3 us per 128-channel tile); 128-bit stores.
// Tiles are 64 consecutive positions of one image (H*W % 64 == 0).  One red.global.add flush per CTA, then the
// fused peer all-reduce tail (ctvq_peer.cuh) like every other backward kernel.
#include <string.h>

#include "ctvq_tc_ptx.cuh"

namespace ctvq {
namespace {

constexpr int kRingThreads = 512;
constexpr int kNW = 8;             // worker warps
constexpr int kNAcc = 8;           // acc warps (4: the workers waited on them a quarter of the time, round-2 ncu)

__device__ __forceinline__ void nb_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {  // hinted wait, ~2 s bound then trap
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
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// D: channels (32 / 64 / 128); TMR: rows per tile (D * TMR = 4096 elements: 128 / 64 / 32); NST: g_out ring depth.
// K and H*W are run-time (K * D floats of accumulator, H*W % TMR == 0).
// seg_shift: log2(tiles per image) when that is a power of two, else -1.
template <int D, int TMR, int NST>
__global__ void __launch_bounds__(kRingThreads, 1) vq_bwd_c1_ring_kernel(const BwdParams p, const int ntiles, const int seg_shift,
                                                                          const __grid_constant__ CUtensorMap gomap,
                                                                          const __grid_constant__ CUtensorMap zmap) {
    constexpr int kTMr = TMR;            // rows per tile
    constexpr int kZS = TMR + 2;         // padded row stride (8-byte aligned rows for 8-byte cp.async; 2-way bank conflicts with lanes along channels)
    constexpr int RB = TMR > 64 ? 7 : 6; // row bits of a list entry (code << RB | row, 16 bits)
    constexpr unsigned RM = (1u << RB) - 1u;
    constexpr int JCH = D / 32;          // 32-channel chunks
    constexpr int R = kNAcc / JCH;       // code residue classes per chunk (acc warp = (chunk, k mod R))
    constexpr int GOF = D * kTMr;        // floats per g_out stage
    constexpr int NZB = 3;               // z / index buffers
    constexpr int kDiff = 1, kAccDone = 4, kWork = 7;  // named barriers: kDiff + b, kAccDone + b (b = buffer), kWork
    constexpr int kBoth = (kNW + kNAcc) * 32;
    static_assert((D == 32 || D == 64 || D == 128) && D * TMR == 4096, "16 KB tiles");
    static_assert((R & (R - 1)) == 0 && R >= 1, "residue classes");
    extern __shared__ __align__(128) float smem[];
    const int K = p.K, HW = p.HW;
    const int KD = K * D;
    float* go_s = smem;                                              // [NST][D][64]
    float* zs = go_s + NST * GOF;                                    // [NZB][D][66]: z, then q - z in place
    int* idx_s = reinterpret_cast<int*>(zs + NZB * D * kZS);         // [NZB][64] range-checked indices
    unsigned short* list_s = reinterpret_cast<unsigned short*>(idx_s + NZB * kTMr);  // [kNAcc][64] per-acc-warp row lists
    uint64_t* bars = reinterpret_cast<uint64_t*>(list_s + kNAcc * kTMr);  // full[NST], empty[NST], zfull[NZB]
    float* acc = reinterpret_cast<float*>(bars + 2 * NST + NZB + 1);  // [K][D]
    const uint32_t bar_full = s_u32(&bars[0]), bar_empty = s_u32(&bars[NST]), bar_zfull = s_u32(&bars[2 * NST]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mb_init(bar_full + 8 * i, 1); mb_init(bar_empty + 8 * i, kNW); }
        for (int i = 0; i < NZB; ++i) mb_init(bar_zfull + 8 * i, kNW * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < KD / 4; i += kRingThreads) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float gl = __ldg(p.g_loss);
    const double nd = (double)p.N * (double)D;
    const float coef_e = (float)(2.0 / nd) * gl;
    const float coef_z = (float)(2.0 * (double)p.beta / nd) * gl;
    const float* __restrict__ E = p.E[0];
    __syncthreads();
    const int niter = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool has_go = p.g_out != nullptr;
    const unsigned seg = (unsigned)(HW / kTMr);  // tiles per image
    auto tile_base = [&](int it, unsigned& b, int& r0) {  // 32-bit: ntiles < 2^31
        const unsigned t = blockIdx.x + (unsigned)it * gridDim.x;
        b = seg_shift >= 0 ? (t >> seg_shift) : (t / seg);
        r0 = (int)(t - b * seg) * kTMr;
    };

    if (warp >= kNW) {
        // =========================== acc warps: (chunk, residue class) of the accumulator ===========================
        const int aw = warp - kNW, jc = aw / R, res = aw % R;
        float* ac = acc + jc * 32 + lane;
        unsigned short* lst = list_s + aw * kTMr;  // this warp's rows of the tile, compacted: (code << 6) | row
        const unsigned lt = (1u << lane) - 1u;
        for (int it = 0; it < niter; ++it) {
            const int zb = it % NZB;
            nb_sync(kDiff + zb, kBoth);  // q - z of this tile is in place
            const float* dcol = zs + (size_t)zb * D * kZS + (jc * 32 + lane) * kZS;
            const int* ks = idx_s + zb * kTMr;
            // ---- the tile's rows whose code falls in this warp's residue class, compacted into a list (ballot + prefix
            // popcount, no serial find-first-set chain): the update loop below is warp-uniform, runs over this class's rows
            // only, and reads four (code, row) pairs with ONE broadcast 64-bit load
            int n = 0;
#pragma unroll
            for (int hf = 0; hf < TMR / 32; ++hf) {
                const int kmine = ks[32 * hf + lane];
                const bool mine = (kmine & (R - 1)) == res;
                const unsigned mask = __ballot_sync(0xffffffffu, mine);
                if (mine) lst[n + __popc(mask & lt)] = (unsigned short)((kmine << RB) | (32 * hf + lane));
                n += __popc(mask);
            }
            __syncwarp();
            struct Group { int k[4]; float d[4]; };
            auto unpack = [&](const uint2 pk, Group& g) {  // four (code, row) entries: codes and q - z of the lane's channel
                const unsigned e0 = pk.x & 0xffffu, e1 = pk.x >> 16, e2 = pk.y & 0xffffu, e3 = pk.y >> 16;
                g.k[0] = (int)(e0 >> RB); g.k[1] = (int)(e1 >> RB); g.k[2] = (int)(e2 >> RB); g.k[3] = (int)(e3 >> RB);
                g.d[0] = dcol[e0 & RM]; g.d[1] = dcol[e1 & RM]; g.d[2] = dcol[e2 & RM]; g.d[3] = dcol[e3 & RM];
            };
            auto rmw = [&](const Group& g) {
                const bool distinct = g.k[0] != g.k[1] && g.k[0] != g.k[2] && g.k[0] != g.k[3] && g.k[1] != g.k[2] &&
                                      g.k[1] != g.k[3] && g.k[2] != g.k[3];
                if (distinct) {  // four different words: loads first, then stores
                    const float a0 = ac[g.k[0] * D], a1 = ac[g.k[1] * D], a2 = ac[g.k[2] * D], a3 = ac[g.k[3] * D];
                    ac[g.k[0] * D] = a0 + g.d[0]; ac[g.k[1] * D] = a1 + g.d[1]; ac[g.k[2] * D] = a2 + g.d[2]; ac[g.k[3] * D] = a3 + g.d[3];
                } else {
                    ac[g.k[0] * D] += g.d[0]; ac[g.k[1] * D] += g.d[1]; ac[g.k[2] * D] += g.d[2]; ac[g.k[3] * D] += g.d[3];
                }
            };
            int i = 0;
            if (n >= 4) {
                // software pipeline, two deep: the list entry of group g+2 and the q - z loads of group g+1 fly under the
                // read-modify-write of group g (every shared-memory round trip but the accumulator's own is covered)
                const uint2* l2 = reinterpret_cast<const uint2*>(lst);
                const int ng = n >> 2;
                Group cur, nxt;
                uint2 pk = l2[ng > 1 ? 1 : 0];
                unpack(l2[0], cur);
                for (int g = 1; g < ng; ++g) {
                    const uint2 pk2 = l2[g + 1 < ng ? g + 1 : g];
                    unpack(pk, nxt);
                    rmw(cur);
                    cur = nxt;
                    pk = pk2;
                }
                rmw(cur);
                i = ng << 2;
            }
            for (; i < n; ++i) {  // 0..3 left-over rows
                const unsigned e = lst[i];
                ac[(int)(e >> RB) * D] += dcol[e & RM];
            }
            __syncwarp();  // the list is rewritten for the next tile
            nb_arrive(kAccDone + zb, kBoth);
        }
    } else {
        // =========================== worker warps: staging, q - z, grad_z ===========================================
        const int gt = tid;                      // 0..255
        constexpr int LQ = TMR / 4;              // grad_z pass: lanes per channel row (4 rows each), CPW channels per warp pass
        constexpr int CPW = 32 / LQ;
        const int hsel = lane / LQ;              // which channel of the warp's CPW
        const int m = (lane % LQ) * 4;           // rows m..m+3 in the grad_z pass
        constexpr int RPW = TMR / kNW;           // rows per worker warp in the q - z pass
        constexpr int PP = TMR / 2;              // 8-byte row pairs per channel
        constexpr int CST = kNW * 32 / PP;       // channels covered by one pass of the 256 staging threads
        constexpr int CPT = D / CST;             // 8-byte copies per thread per tile
        // thread gt copies row pair gt % PP of channels gt / PP + CST i: coalesced runs along H*W
        const uint32_t st_dst = s_u32(zs + (gt / PP) * kZS + 2 * (gt % PP));
        const size_t st_src = (size_t)(gt / PP) * HW + 2 * (gt % PP);
        long long kreg = 0;                      // index of row gt (gt < 64) of the most recently staged tile
        auto stage = [&](int it) {               // tile it -> buffer it % NZB (asynchronous; completion arrives on zfull)
            unsigned b;
            int r0;
            tile_base(it, b, r0);
            const int zb = it % NZB;
            const float* src = p.z + (size_t)b * D * HW + r0 + st_src;
            const uint32_t dst = st_dst + (uint32_t)zb * D * kZS * 4u;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (uint32_t)i * CST * kZS * 4u), "l"(src) : "memory");
                src += (size_t)CST * HW;
            }
            if (gt < kTMr) kreg = __ldg(p.idx + (size_t)b * HW + r0 + gt);
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_zfull + 8 * zb) : "memory");
        };
        auto publish_idx = [&](int it) {         // the index loaded by stage(it): range-check, clamp, hand to everybody
            if (gt < kTMr) {
                long long kk = kreg;
                if (kk < 0 || kk >= K) { atomicOr(p.err, 1u); kk = kk < 0 ? 0 : K - 1; }  // caller-supplied index out of range
                idx_s[(it % NZB) * kTMr + gt] = (int)kk;
            }
        };
        auto issue_go = [&](int it) {            // warp 0: stream tile it's g_out block [D][64] into ring slot it % NST
            const int st = it % NST;
            if (it >= NST) mb_wait(bar_empty + 8 * st, (uint32_t)((it / NST) - 1) & 1u);
            unsigned b;
            int r0;
            tile_base(it, b, r0);
            if (lane == 0) {  // ONE tensor-map box per tile (a 256-byte bulk copy per channel costs ~46 cycles of TMA service each)
                mb_expect_tx(bar_full + 8 * st, GOF * 4u);
                tc::tma_load_3d(s_u32(go_s + st * GOF), &gomap, bar_full + 8 * st, r0, 0, (int)b);
            }
            __syncwarp();
        };
        // this warp's RPW rows of tile `it`: wait for the staged tile, read the indices, issue the codebook reads -- RPW x D/32
        // (= 16) coalesced 128-byte rows from L2 per lane.  Called one tile AHEAD, so the L2 latency runs under the grad_z pass of
        // the previous tile instead of stalling every warp at the top of the loop.
        float e[RPW][JCH];
        auto fetch_e = [&](int it) {
            const int zb = it % NZB;
            mb_wait(bar_zfull + 8 * zb, (uint32_t)(it / NZB) & 1u);
            const int* ks = idx_s + zb * kTMr;
            int k[RPW];
#pragma unroll
            for (int u = 0; u < RPW; ++u) k[u] = ks[warp + kNW * u];  // broadcast reads
#pragma unroll
            for (int u = 0; u < RPW; ++u)
#pragma unroll
                for (int c = 0; c < JCH; ++c) e[u][c] = __ldg(E + (size_t)k[u] * D + c * 32 + lane);
        };
        // HBM latency is hidden TWICE: every z tile is requested into L2 kL2 tiles ahead by ONE tensor-map prefetch (no
        // shared-memory destination, so it costs no buffer), and staged from L2 two tiles ahead -- with three buffers the
        // staging lead alone is one iteration, less than an HBM round trip under load
        constexpr int kL2 = 5;
        auto prefetch_z = [&](int it) {
            unsigned b;
            int r0;
            tile_base(it, b, r0);
            asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(&zmap), "r"(r0), "r"(0), "r"((int)b) : "memory");
            if (has_go)  // the g_out ring runs NST-1 tiles ahead: same double cover
                asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(&gomap), "r"(r0), "r"(0), "r"((int)b) : "memory");
        };
        if (tid == 0)
            for (int it = 2; it < kL2 && it < niter; ++it) prefetch_z(it);
        if (niter > 0) { stage(0); publish_idx(0); }
        if (niter > 1) { stage(1); publish_idx(1); }
        if (warp == 0 && has_go)
            for (int it = 0; it < NST - 1 && it < niter; ++it) issue_go(it);
        nb_sync(kWork, kNW * 32);                // the first tiles' indices are published
        if (niter > 0) fetch_e(0);
        for (int it = 0; it < niter; ++it) {
            const int zb = it % NZB, st = it % NST;
            if (warp == 0 && has_go && it + NST - 1 < niter) issue_go(it + NST - 1);
            if (tid == 0 && it + kL2 < niter) prefetch_z(it + kL2);
            if (it >= 1 && it + 1 < niter) publish_idx(it + 1);  // loaded by the refill of the previous iteration
            // ---- q - z in place (the codebook values were requested one tile ago) ---------------------------------------
            {
                float* zt = zs + (size_t)zb * D * kZS;
#pragma unroll
                for (int u = 0; u < RPW; ++u)
#pragma unroll
                    for (int c = 0; c < JCH; ++c) {
                        float* zp = zt + (c * 32 + lane) * kZS + warp + kNW * u;
                        *zp = __fsub_rn(e[u][c], *zp);  // q - z
                    }
            }
            nb_arrive(kDiff + zb, kBoth);        // the acc warps may start on this tile
            nb_sync(kWork, kNW * 32);            // every worker's share of q - z is in place (and grad_z of tile it-1 is done)
            // ---- refill: tile it+2 re-uses the buffer of tile it-1 (its grad_z is done; wait for its accumulation) ----
            if (it >= 1) nb_sync(kAccDone + (it + 2) % NZB, kBoth);
            if (it + 2 < niter) stage(it + 2);
            if (it + 1 < niter) fetch_e(it + 1);  // next tile's codebook rows fly under this tile's grad_z
            // ---- grad_z = g_out - coef_z (q - z): lanes along H*W, 128-bit ------------------------------------------------
            unsigned b;
            int r0;
            tile_base(it, b, r0);
            const float* zt = zs + (size_t)zb * D * kZS;
            const float* gos = go_s + st * GOF;
            float* gz_row = p.gz + (size_t)b * D * HW + r0 + m;
            if (has_go) mb_wait(bar_full + 8 * st, (uint32_t)(it / NST) & 1u);
#pragma unroll 4
            for (int ch = CPW * warp + hsel; ch < D; ch += CPW * kNW) {
                const float2 da = *reinterpret_cast<const float2*>(zt + ch * kZS + m), db = *reinterpret_cast<const float2*>(zt + ch * kZS + m + 2);
                float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_go) go = *reinterpret_cast<const float4*>(gos + ch * kTMr + m);
                *reinterpret_cast<float4*>(gz_row + (size_t)ch * HW) =
                    make_float4(go.x - coef_z * da.x, go.y - coef_z * da.y, go.z - coef_z * db.x, go.w - coef_z * db.y);
            }
            __syncwarp();
            if (lane == 0 && has_go) mb_arrive(bar_empty + 8 * st);  // ring slot may be refilled
        }
    }
    __syncthreads();
    for (int i = tid; i < KD; i += kRingThreads) {
        const float v = acc[i];
        if (v != 0.0f) atomicAdd(&p.gE[i], coef_e * v);
    }
    peer_tail(p.peer, p.gE);  // fused collective (no-op unless ctvq_backward_allreduce armed it)
}

template <int D, int TMR>
size_t ring_smem(int K, int nst) {
    return ((size_t)nst * D * TMR + 3 * (size_t)D * (TMR + 2)) * 4 + 3 * TMR * 4 + kNAcc * TMR * 2 + (2 * (size_t)nst + 3 + 1) * 8 + (size_t)K * D * 4;
}

template <int D, int TMR, int NST>
int launch_ring(const BwdParams& p, cudaStream_t s) {
    const size_t smem = ring_smem<D, TMR>(p.K, NST);
    const long long nt = p.N / TMR;
    if (nt > 0x7fffffffLL || p.B > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
    int grid = sm_count();
    if (grid > nt) grid = (int)nt;
    CUtensorMap gomap, zmap;
    memset(&gomap, 0, sizeof(gomap));
    if (p.g_out != nullptr && tc::make_plain_map(gomap, p.g_out, CTVQ_F32, p.HW, D, p.B, TMR, D) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    if (tc::make_plain_map(zmap, p.z, CTVQ_F32, p.HW, D, p.B, TMR, D) != CTVQ_OK) return CTVQ_E_UNSUPPORTED;
    auto kern = vq_bwd_c1_ring_kernel<D, TMR, NST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int seg = p.HW / TMR;
    int seg_shift = -1;
    if ((seg & (seg - 1)) == 0)
        for (seg_shift = 0; (1 << seg_shift) < seg; ++seg_shift) {}
    kern<<<grid, kRingThreads, smem, s>>>(p, (int)nt, seg_shift, gomap, zmap);
    return (int)cudaGetLastError();
}

template <int D, int TMR>
int launch_ring_nst(const BwdParams& p, cudaStream_t s) {
    const size_t cap = 227 * 1024 - 256;  // static shared memory of the tail + slack
    if (p.HW % TMR != 0 || p.N < (long long)sm_count() * TMR * 4) return CTVQ_E_UNSUPPORTED;
    if (ring_smem<D, TMR>(p.K, 3) <= cap) return launch_ring<D, TMR, 3>(p, s);
    if (ring_smem<D, TMR>(p.K, 2) <= cap) return launch_ring<D, TMR, 2>(p, s);
    return CTVQ_E_UNSUPPORTED;
}
}  // namespace

// CTVQ_E_UNSUPPORTED: not a single full-width fp32 codebook of 32 / 64 channels, H*W not a multiple of 64, unaligned
// tensors, an accumulator beyond shared memory, or a batch too small to amortise zeroing + flushing it per CTA.
int launch_backward_ring(const BwdParams& p, cudaStream_t s) {
    if (p.dtype != CTVQ_F32 || p.C != 1 || p.d != p.Dtot) return CTVQ_E_UNSUPPORTED;
    // the 16-bit list entries hold the code next to 6 (7 at 128-row tiles) row bits
    if (p.K < 128 || p.K > (p.d == 32 ? 512 : 1024) || (p.K * p.d) % 4 != 0) return CTVQ_E_UNSUPPORTED;  // few codes: the ownership kernels of ctvq_bwd_fast.cu / the tiled kernel
    if ((reinterpret_cast<uintptr_t>(p.z) & 15) || (reinterpret_cast<uintptr_t>(p.gz) & 15) ||
        (p.g_out && (reinterpret_cast<uintptr_t>(p.g_out) & 15)) || (reinterpret_cast<uintptr_t>(p.idx) & 7))
        return CTVQ_E_UNSUPPORTED;
    // every CTA zeroes and flushes a [K, d] accumulator: only worth it when the rows outweigh that
    if (p.N < (long long)sm_count() * p.K) return CTVQ_E_UNSUPPORTED;
    if (p.d == 64) return launch_ring_nst<64, 64>(p, s);
    if (p.d == 128) return launch_ring_nst<128, 32>(p, s);
    if (p.d == 32) return launch_ring_nst<32, 128>(p, s);
    return CTVQ_E_UNSUPPORTED;
}

}  // namespace ctvq
