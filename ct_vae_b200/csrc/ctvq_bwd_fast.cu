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
#include "ctvq_common.cuh"

namespace ctvq {
namespace {

constexpr int kBT = 256;
constexpr int kTM = 128;

template <int D, int C, int K, int HWT, int DTOT, int CS>
__global__ void __launch_bounds__(kBT, 2) vq_bwd_fast_kernel(const BwdParams p, const int ntiles) {
    constexpr int USED = (C - 1) * CS + D;
    constexpr int ZS = kTM + 1;
    constexpr int ESD = D + 1;
    constexpr int CKD = C * K * D;
    constexpr int JCH = (D + 31) / 32;
    constexpr int ITEMS = C * JCH;
    static_assert(ITEMS <= 4, "one accumulation warp per (codebook, channel chunk)");
    static_assert(HWT % 4 == 0 && kTM % 4 == 0, "row quads");
    extern __shared__ __align__(16) float smem[];
    int* idx_s = reinterpret_cast<int*>(smem);  // [C][TM]
    float* zs = smem + C * kTM;                 // [USED][ZS]
    float* acc = zs + USED * ZS;                // [C][K][D]
    float* es = acc + CKD;                      // [C][K][D+1]

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

    // ---- register prefetch of a tile's staging data (indices + touched z channels), one tile ahead ---------------
    constexpr int NZ = (USED + 1) / 2;        // z channels per thread (thread = row m, channel parity tid/128)
    constexpr int NI = (C + 1) / 2;
    float zreg[NZ];
    long long kreg[NI];  // raw int64 indices: validated when they are stored to shared memory, not when loaded
    const int sm_m = tid & (kTM - 1), sm_par = tid / kTM;
    auto prefetch = [&](int tile) {
        const long long row0 = (long long)tile * kTM;
        const long long n = row0 + sm_m;
        const bool valid = n < p.N;
        const long long b = valid ? n / HWT : 0;
        const int hw = valid ? (int)(n - b * HWT) : 0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int c = sm_par + 2 * i;
            kreg[i] = (valid && c < C) ? __ldg(p.idx + ((size_t)b * C + c) * HWT + hw) : 0;
        }
        const float* src = p.z + (size_t)b * DTOT * HWT + hw;
#pragma unroll
        for (int i = 0; i < NZ; ++i) {
            const int ch = sm_par + 2 * i;
            zreg[i] = (valid && ch < USED) ? __ldg(src + (size_t)ch * HWT) : 0.0f;
        }
    };
    if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * kTM;
        const int mcount = (int)min((long long)kTM, p.N - row0);
        __syncthreads();  // previous tile fully consumed
#pragma unroll
        for (int i = 0; i < NI; ++i)
            if (sm_par + 2 * i < C) {
                long long kk = kreg[i];
                if (kk < 0 || kk >= K) { atomicOr(p.err, 1u); kk = kk < 0 ? 0 : K - 1; }
                idx_s[(sm_par + 2 * i) * kTM + sm_m] = (int)kk;
            }
#pragma unroll
        for (int i = 0; i < NZ; ++i)
            if (sm_par + 2 * i < USED) zs[(sm_par + 2 * i) * ZS + sm_m] = zreg[i];
        __syncthreads();
        if (tile + (int)gridDim.x < ntiles) {
            prefetch(tile + gridDim.x);  // latency hidden behind this tile's work
            // pull the NEXT tile's g_out rows (contiguous: whole images) into L2 so the pipelined 128-bit loads of
            // the grad_z warps see L2 latency instead of HBM latency
            if (tid == 0 && p.g_out && (kTM % HWT) == 0) {
                const long long nrow0 = (long long)(tile + gridDim.x) * kTM;
                const long long nrows = min((long long)kTM, p.N - nrow0);
                const float* src = p.g_out + (size_t)(nrow0 / HWT) * C * D * HWT;
                const unsigned bytes = (unsigned)((nrows / HWT) * C * D * HWT * sizeof(float));
                if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
            }
        }
        if (warp >= 4) {
            // ---- grad_z: warp handles channels ch = (warp-4) + 4*i; lane handles rows 4*lane .. 4*lane+3 -------------
            const int m = lane * 4;
            if (m < mcount) {
                const long long n = row0 + m;
                const long long b = n / HWT;
                const int hw = (int)(n - b * HWT);
                const float* go_row = p.g_out ? p.g_out + (size_t)b * C * D * HWT + hw : nullptr;
                float* gz_row = p.gz + (size_t)b * DTOT * HWT + hw;
                int eb[C][4];  // shared-memory float index of this row's codeword per codebook
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int4 kk = *reinterpret_cast<const int4*>(idx_s + c * kTM + m);
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
                // channels that no slice reads: zeros, no loads (issued first so the stores drain under the loads)
                for (int ch = USED + ((warp - 4) - USED % 4 + 4) % 4; ch < DTOT; ch += 4)
                    *reinterpret_cast<float4*>(gz_row + (size_t)ch * HWT) = make_float4(0.f, 0.f, 0.f, 0.f);
                // active channels, software-pipelined: the next channel's g_out loads fly while this one is computed
                float4 go_cur[C], go_nxt[C];
                int ch = warp - 4;
                load_go(ch, go_cur);
                for (; ch < USED; ch += 4) {
                    load_go(ch + 4, go_nxt);
                    const float z0 = zs[ch * ZS + m], z1 = zs[ch * ZS + m + 1], z2 = zs[ch * ZS + m + 2], z3 = zs[ch * ZS + m + 3];
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = ch - c * CS;
                        if (j >= 0 && j < D) {
                            const float d0 = __fsub_rn(es[eb[c][0] + j], z0), d1 = __fsub_rn(es[eb[c][1] + j], z1);
                            const float d2 = __fsub_rn(es[eb[c][2] + j], z2), d3 = __fsub_rn(es[eb[c][3] + j], z3);
                            g.x += go_cur[c].x - coef_z * d0; g.y += go_cur[c].y - coef_z * d1;
                            g.z += go_cur[c].z - coef_z * d2; g.w += go_cur[c].w - coef_z * d3;
                        }
                    }
                    *reinterpret_cast<float4*>(gz_row + (size_t)ch * HWT) = g;
#pragma unroll
                    for (int c = 0; c < C; ++c) go_cur[c] = go_nxt[c];
                }
            }
        } else if (warp < ITEMS) {
            // ---- codebook gradient: this warp owns accumulator slab (c, 32-channel chunk) -----------------------------------
            const int c = warp / JCH;
            const int j = (warp - c * JCH) * 32 + lane;
            const bool act = j < D;
            const int jj = act ? j : 0;
            const float* zcol = zs + (c * CS + jj) * ZS;
            const int* ks = idx_s + c * kTM;
            float* ac = acc + c * K * D + jj;
            const float* ec = es + c * K * ESD + jj;
            int m = 0;
            for (; m + 4 <= mcount; m += 4) {
                const int4 kk = *reinterpret_cast<const int4*>(ks + m);
                const bool distinct = kk.x != kk.y && kk.x != kk.z && kk.x != kk.w && kk.y != kk.z && kk.y != kk.w &&
                                      kk.z != kk.w;
                if (act) {
                    const float d0 = __fsub_rn(ec[kk.x * ESD], zcol[m]), d1 = __fsub_rn(ec[kk.y * ESD], zcol[m + 1]);
                    const float d2 = __fsub_rn(ec[kk.z * ESD], zcol[m + 2]), d3 = __fsub_rn(ec[kk.w * ESD], zcol[m + 3]);
                    if (distinct) {
                        const float a0 = ac[kk.x * D], a1 = ac[kk.y * D], a2 = ac[kk.z * D], a3 = ac[kk.w * D];
                        ac[kk.x * D] = a0 + d0; ac[kk.y * D] = a1 + d1; ac[kk.z * D] = a2 + d2; ac[kk.w * D] = a3 + d3;
                    } else {
                        ac[kk.x * D] += d0; ac[kk.y * D] += d1; ac[kk.z * D] += d2; ac[kk.w * D] += d3;
                    }
                }
            }
            for (; m < mcount; ++m)
                if (act) ac[ks[m] * D] += __fsub_rn(ec[ks[m] * ESD], zcol[m]);
        }
    }
    __syncthreads();
    for (int i = tid; i < CKD; i += kBT) {
        const float v = acc[i];
        if (v != 0.0f) atomicAdd(&p.gE[i], coef_e * v);
    }
}

template <int D, int C, int K, int HWT, int DTOT, int CS>
int launch(const BwdParams& p, cudaStream_t s) {
    constexpr int USED = (C - 1) * CS + D;
    constexpr size_t smem = sizeof(float) * ((size_t)C * kTM + (size_t)USED * (kTM + 1) + (size_t)C * K * D + (size_t)C * K * (D + 1));
    static_assert(smem <= 113 * 1024, "two CTAs per SM");
    const long long nt = (p.N + kTM - 1) / kTM;
    if (nt > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
    int grid = 148 * 2;
    if ((long long)grid * 4 > nt) grid = (int)((nt + 3) / 4);  // small problems: fewer accumulator flushes
    if (grid < 1) grid = 1;
    auto kern = vq_bwd_fast_kernel<D, C, K, HWT, DTOT, CS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kBT, smem, s>>>(p, (int)nt);
    return (int)cudaGetLastError();
}
}  // namespace

int launch_backward_fast(const BwdParams& p, cudaStream_t s) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(p.gz) & 15) == 0) &&
                         (p.g_out == nullptr || (reinterpret_cast<uintptr_t>(p.g_out) & 15) == 0);
    if (!aligned) return CTVQ_E_UNSUPPORTED;
    // configs/mcq_vae.yaml: C=4, d=32, K=64, latents [B,128,8,8], overlapping slices
    if (p.d == 32 && p.C == 4 && p.K == 64 && p.HW == 64 && p.Dtot == 128 && p.cs == 1) return launch<32, 4, 64, 64, 128, 1>(p, s);
    return CTVQ_E_UNSUPPORTED;
}

}  // namespace ctvq
