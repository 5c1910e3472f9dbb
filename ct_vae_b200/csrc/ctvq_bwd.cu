// Tiled backward kernel: straight-through + commitment gradient (coalesced, 128-bit) and the codebook
// gradient accumulated WITHOUT atomics in the inner loop.
//
// Replaces autograd through models/vq_vae.py:43-53 for all C codebooks of models/mcq_vae.py:112-127 in one
// launch.  Per tile of TM latent rows a CTA
//   phase 0  stages the indices and the z channels the slices touch in shared memory (coalesced along HW),
//   phase 1  lanes along HW: gz[b,ch,p] = sum_{c: ch in slice c} g_out[b,c*d+j,p] - coef_z*(E_c[k][j] - z)
//            written with 128-bit stores (untouched channels get zeros),
//   phase 2  lanes along the channel j: each warp OWNS a (codebook, 32-channel chunk) slab of the shared
//            [C,K,d] accumulator and walks the tile's rows, acc[k][j] += E_c[k][j] - z — plain shared-memory
//            read-modify-write, race-free by ownership, 4 rows in flight when their codes are distinct.
// The accumulator is flushed once per CTA (persistent grid) with red.global.add.f32.
#include "ctvq_common.cuh"

namespace ctvq {

namespace {
constexpr int kBT = 256;  // threads

template <int TM, int VEC, typename T>
__global__ void __launch_bounds__(kBT) vq_bwd_tile_kernel(const BwdParams p, const int use_es, const int ntiles) {
    const T* __restrict__ zT = reinterpret_cast<const T*>(p.z);
    const T* __restrict__ goT = reinterpret_cast<const T*>(p.g_out);
    T* __restrict__ gzT = reinterpret_cast<T*>(p.gz);
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, d = p.d, K = p.K, HW = p.HW, Dtot = p.Dtot, cs = p.cs;
    const int used = min(Dtot, (C - 1) * cs + d);
    constexpr int ZS = TM + 1;
    const int ESd = d + 1;
    const int ckd = C * K * d;
    int* idx_s = reinterpret_cast<int*>(smem);          // [C][TM]  (first: keeps the int4 reads 16B aligned)
    float* zs = smem + C * TM;                          // [used][ZS]
    float* acc = zs + used * ZS;                        // [C][K][d]
    float* es = acc + ckd;                              // [C][K][d+1] when use_es

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < ckd; i += kBT) acc[i] = 0.0f;
    if (use_es) {
        for (int i = tid; i < ckd; i += kBT) {
            const int j = i % d, ck = i / d;
            es[ck * ESd + j] = IO<T>::cb(__ldg(p.E[ck / K] + (size_t)(ck % K) * d + j));
        }
    }
    const float gl = __ldg(p.g_loss);
    const double nd = (double)p.N * (double)d;
    const float coef_e = (float)(2.0 / nd) * gl;                    // weight of (q - z) in d vq_loss / d E
    const float coef_z = (float)(2.0 * (double)p.beta / nd) * gl;   // weight of (z - q) in d vq_loss / d z
    const int jchunks = (d + 31) >> 5;
    const int items = C * jchunks;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * TM;
        const int mcount = (int)min((long long)TM, p.N - row0);
        __syncthreads();  // previous tile fully consumed (and acc/es initialised on the first pass)
        // ---- phase 0: stage indices and z ------------------------------------------------------------------
        {
            const int m = tid % TM;
            const bool valid = m < mcount;
            const long long n = row0 + m;
            const long long b = valid ? n / HW : 0;
            const int pp = valid ? (int)(n - b * HW) : 0;
            for (int c = tid / TM; c < C; c += kBT / TM) {
                int k = 0;
                if (valid) {
                    const long long kk = __ldg(p.idx + ((size_t)b * C + c) * HW + pp);
                    if (kk < 0 || kk >= K) { atomicOr(p.err, 1u); k = kk < 0 ? 0 : K - 1; } else k = (int)kk;
                }
                idx_s[c * TM + m] = k;
            }
            const T* src = zT + (size_t)b * Dtot * HW + pp;
            for (int ch = tid / TM; ch < used; ch += kBT / TM)
                zs[ch * ZS + m] = valid ? IO<T>::ld(src + (size_t)ch * HW) : 0.0f;
        }
        __syncthreads();
        // ---- phase 1: gz, lanes along HW ------------------------------------------------------------------
        constexpr int MG = TM / VEC;  // row groups per channel
        for (int it = tid; it < Dtot * MG; it += kBT) {
            const int ch = it / MG, m = (it % MG) * VEC;
            if (m >= mcount) continue;
            const long long n = row0 + m;
            const long long b = n / HW;
            const int pp = (int)(n - b * HW);
            float g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g[u] = 0.0f;
            int c_hi, c_lo = 0;
            if (cs > 0) {
                c_hi = min(ch / cs, C - 1);
                if (ch - d + 1 > 0) c_lo = (ch - d + cs) / cs;
            } else {
                c_hi = (ch < d) ? C - 1 : -1;
            }
            for (int c = c_lo; c <= c_hi; ++c) {
                const int j = ch - c * cs;
                float go[4] = {0.f, 0.f, 0.f, 0.f};
                const size_t goff = ((size_t)b * C * d + (size_t)c * d + j) * HW + pp;
                if (VEC == 4) {
                    if (goT) IO<T>::ld4(goT + goff, go);
                } else {
                    go[0] = goT ? IO<T>::ld(goT + goff) : 0.0f;
                }
                const float* __restrict__ Ec = p.E[c];
#pragma unroll
                for (int u = 0; u < VEC; ++u) {
                    const int k = idx_s[c * TM + m + u];
                    const float e = use_es ? es[(c * K + k) * ESd + j] : IO<T>::cb(__ldg(Ec + (size_t)k * d + j));
                    const float diff = __fsub_rn(e, zs[ch * ZS + m + u]);  // q - z
                    g[u] += go[u] - coef_z * diff;
                }
            }
            T* dst = gzT + ((size_t)b * Dtot + ch) * HW + pp;
            if (VEC == 4) IO<T>::st4(dst, g);
            else IO<T>::st(dst, g[0]);
        }
        // ---- phase 2: codebook-gradient accumulation, lanes along the channel -----------------------------------
        for (int item = warp; item < items; item += kBT / 32) {
            const int c = item / jchunks;
            const int j = (item - c * jchunks) * 32 + lane;
            const bool act = j < d;
            const int jj = act ? j : 0;
            const float* __restrict__ Ec = p.E[c];
            const float* zcol = zs + (c * cs + jj) * ZS;
            const int* ks = idx_s + c * TM;
            float* ac = acc + (size_t)c * K * d + jj;
            const float* ec = es + (size_t)c * K * ESd + jj;
            int m = 0;
            for (; m + 4 <= mcount; m += 4) {
                const int4 kk = *reinterpret_cast<const int4*>(ks + m);
                const bool distinct = kk.x != kk.y && kk.x != kk.z && kk.x != kk.w && kk.y != kk.z && kk.y != kk.w &&
                                      kk.z != kk.w;
                if (act) {
                    const float e0 = use_es ? ec[kk.x * ESd] : IO<T>::cb(__ldg(Ec + (size_t)kk.x * d + jj));
                    const float e1 = use_es ? ec[kk.y * ESd] : IO<T>::cb(__ldg(Ec + (size_t)kk.y * d + jj));
                    const float e2 = use_es ? ec[kk.z * ESd] : IO<T>::cb(__ldg(Ec + (size_t)kk.z * d + jj));
                    const float e3 = use_es ? ec[kk.w * ESd] : IO<T>::cb(__ldg(Ec + (size_t)kk.w * d + jj));
                    const float d0 = __fsub_rn(e0, zcol[m]), d1 = __fsub_rn(e1, zcol[m + 1]);
                    const float d2 = __fsub_rn(e2, zcol[m + 2]), d3 = __fsub_rn(e3, zcol[m + 3]);
                    if (distinct) {
                        const float a0 = ac[kk.x * d], a1 = ac[kk.y * d], a2 = ac[kk.z * d], a3 = ac[kk.w * d];
                        ac[kk.x * d] = a0 + d0; ac[kk.y * d] = a1 + d1; ac[kk.z * d] = a2 + d2; ac[kk.w * d] = a3 + d3;
                    } else {
                        ac[kk.x * d] += d0; ac[kk.y * d] += d1; ac[kk.z * d] += d2; ac[kk.w * d] += d3;
                    }
                }
            }
            for (; m < mcount; ++m) {
                const int k = ks[m];
                if (act) {
                    const float e = use_es ? ec[k * ESd] : IO<T>::cb(__ldg(Ec + (size_t)k * d + jj));
                    ac[k * d] += __fsub_rn(e, zcol[m]);
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < ckd; i += kBT) {
        const float v = acc[i];
        if (v != 0.0f) atomicAdd(&p.gE[i], coef_e * v);
    }
    peer_tail(p.peer, p.gE);  // fused collective (no-op unless ctvq_backward_allreduce armed it)
}
}  // namespace

namespace {
template <typename T>
int launch_tiles(const BwdParams& p, int use_es, int ntiles, int grid, size_t sm, cudaStream_t s);
}

// returns CTVQ_E_UNSUPPORTED when the accumulator does not fit in shared memory (caller falls back)
int launch_backward_tiled(const BwdParams& p, cudaStream_t s) {
    constexpr int TM = 128;
    const int used = p.Dtot < (p.C - 1) * p.cs + p.d ? p.Dtot : (p.C - 1) * p.cs + p.d;
    const size_t ckd = (size_t)p.C * p.K * p.d;
    const size_t base = ((size_t)used * (TM + 1) + ckd + (size_t)p.C * TM + 4) * sizeof(float);
    const size_t es_b = (size_t)p.C * p.K * (p.d + 1) * sizeof(float);
    int use_es, per_sm;
    if (base + es_b <= 113 * 1024) { use_es = 1; per_sm = 2; }
    else if (base + es_b <= 220 * 1024) { use_es = 1; per_sm = 1; }
    else if (base <= 220 * 1024) { use_es = 0; per_sm = base <= 113 * 1024 ? 2 : 1; }
    else return CTVQ_E_UNSUPPORTED;
    const size_t sm = base + (use_es ? es_b : 0);
    const long long ntiles_ll = (p.N + TM - 1) / TM;
    if (ntiles_ll > 0x7fffffffLL) return CTVQ_E_UNSUPPORTED;
    const int ntiles = (int)ntiles_ll;
    int grid = sm_count() * per_sm;
    // small problems: do not pay one accumulator flush per idle CTA
    const int min_tiles_per_cta = 4;
    if ((long long)grid * min_tiles_per_cta > ntiles) grid = (ntiles + min_tiles_per_cta - 1) / min_tiles_per_cta;
    if (grid < 1) grid = 1;
    // tiny batches against a big codebook: zeroing + flushing a [C,K,d] accumulator per CTA costs more than adding the
    // N*C*d differences straight into global memory -> let the caller fall back to the direct-atomic kernel
    if ((double)p.N * p.C * p.d < 2.0 * (double)grid * (double)ckd) return CTVQ_E_UNSUPPORTED;
    return p.dtype == CTVQ_BF16 ? launch_tiles<__nv_bfloat16>(p, use_es, ntiles, grid, sm, s) : launch_tiles<float>(p, use_es, ntiles, grid, sm, s);
}

namespace {
template <typename T>
int launch_tiles(const BwdParams& p, int use_es, int ntiles, int grid, size_t sm, cudaStream_t s) {
    constexpr int TM = 128;
    const bool vec = (p.HW % 4 == 0) && IO<T>::aligned4(p.gz) && (p.g_out == nullptr || IO<T>::aligned4(p.g_out));
    cudaError_t e;
    if (vec) {
        e = cudaFuncSetAttribute(vq_bwd_tile_kernel<TM, 4, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        vq_bwd_tile_kernel<TM, 4, T><<<grid, kBT, sm, s>>>(p, use_es, ntiles);
    } else {
        e = cudaFuncSetAttribute(vq_bwd_tile_kernel<TM, 1, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        vq_bwd_tile_kernel<TM, 1, T><<<grid, kBT, sm, s>>>(p, use_es, ntiles);
    }
    return (int)cudaGetLastError();
}
}  // namespace

}  // namespace ctvq
