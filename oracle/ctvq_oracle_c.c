/* Plain-C restatement of the quantiser arithmetic — TEST INFRASTRUCTURE, NOT A PRODUCT PATH.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * It follows the reference's formulas (models/vq_vae.py:30-35 distances+argmin, :43 gather,
 * :47-53 losses + straight-through, autograd of those for the gradients; models/mcq_vae.py:104,117
 * the overlapping channel slices [:, i:i+d]) but fixes ONE evaluation order for every sum — the
 * order the CUDA kernels are specified to use (DESIGN.md "arithmetic contract"):
 *     |z|^2, |e|^2, z.e :  acc = fmaf(a_j, b_j, acc), j ascending from 0, acc starts at 0
 *     dist  = (zz + ee_k) - 2*dot          two fp32 roundings, the reference's association
 *     argmin: ascending k, strict '<' (first minimum wins); the first NaN wins over any number
 * so the GPU kernels can be held bit-exact against it on EVERY row, while the comparison against
 * the reference proper (ATen sgemm order; oracle/ctvq_oracle.py + tests/golden) allows only
 * counted near-ties.  Pinned by tests/test_oracle_golden.py::test_c_oracle_* against the goldens.
 *
 * Layout: latents NCHW [B, Dtot, HW]; codebook c is [K, d] and reads channels c*cs .. c*cs+d-1;
 * quantised output [B, C*d, HW]; indices int64 [B, C, HW].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float dot_seq(const float *a, long sa, const float *b, long sb, int d) {
    float acc = 0.0f;
    for (int j = 0; j < d; ++j) acc = fmaf(a[j * sa], b[j * sb], acc);
    return acc;
}

/* models/vq_vae.py:30-35 (mcq_vae.py:31-39 per codebook, :104 slice) */
void ctvq_c_argmin(const float *z, const float *const *E, int64_t B, int Dtot, int HW, int C, int d, int K,
                   int cs, int64_t *idx_out) {
    float *ee = (float *)malloc(sizeof(float) * (size_t)K);
    for (int c = 0; c < C; ++c) {
        const float *Ec = E[c];
        for (int k = 0; k < K; ++k) ee[k] = dot_seq(Ec + (long)k * d, 1, Ec + (long)k * d, 1, d);
        for (int64_t b = 0; b < B; ++b)
            for (int p = 0; p < HW; ++p) {
                const float *zr = z + ((b * Dtot + (long)c * cs) * HW + p);
                float zz = dot_seq(zr, HW, zr, HW, d);
                float best = 0.0f;
                int bi = -1;
                for (int k = 0; k < K; ++k) {
                    float dot = dot_seq(zr, HW, Ec + (long)k * d, 1, d);
                    float dist = (zz + ee[k]) - 2.0f * dot;
                    int take = (bi < 0) || (dist < best) || (dist != dist && best == best);
                    if (take) { best = dist; bi = k; }
                }
                idx_out[(b * C + c) * HW + p] = bi;
            }
    }
    free(ee);
}

/* Near-tie accounting (BASELINE.json north_star; include/ctvq.h CTVQ_NEAR_TIE_REL): number of (row, codebook) pairs
 * whose best and second-best fp32 distances (models/vq_vae.py:30-32, evaluation order above) d1 <= d2 satisfy
 * d2 - d1 <= 1e-6 * |d1| in fp32 arithmetic.  Rows with a non-finite distance are not counted. */
int64_t ctvq_c_neartie_count(const float *z, const float *const *E, int64_t B, int Dtot, int HW, int C, int d, int K,
                             int cs) {
    float *ee = (float *)malloc(sizeof(float) * (size_t)K);
    int64_t count = 0;
    for (int c = 0; c < C; ++c) {
        const float *Ec = E[c];
        for (int k = 0; k < K; ++k) ee[k] = dot_seq(Ec + (long)k * d, 1, Ec + (long)k * d, 1, d);
        for (int64_t b = 0; b < B; ++b)
            for (int p = 0; p < HW; ++p) {
                const float *zr = z + ((b * Dtot + (long)c * cs) * HW + p);
                float zz = dot_seq(zr, HW, zr, HW, d);
                float d1 = INFINITY, d2 = INFINITY;
                int finite = 1;
                for (int k = 0; k < K; ++k) {
                    float dot = dot_seq(zr, HW, Ec + (long)k * d, 1, d);
                    float dist = (zz + ee[k]) - 2.0f * dot;
                    if (!isfinite(dist)) { finite = 0; break; }
                    if (dist < d1) { d2 = d1; d1 = dist; }
                    else if (dist < d2) d2 = dist;
                }
                if (finite && K >= 2) {
                    volatile float gap = d2 - d1;
                    volatile float lim = 1e-6f * fabsf(d1);
                    if (gap <= lim) ++count;  /* '<=': an exact tie at distance 0 counts too */
                }
            }
    }
    free(ee);
    return count;
}

/* models/vq_vae.py:43-55 (mcq_vae.py:45-64 per codebook, :117-125 slice/cat/sum) */
void ctvq_c_gather_st_loss(const float *z, const float *const *E, const int64_t *idx, int64_t B, int Dtot, int HW,
                           int C, int d, int K, int cs, float beta, float *q_out, float *loss_out /* [C+1] */) {
    (void)K;
    float total = 0.0f;
    for (int c = 0; c < C; ++c) {
        double acc = 0.0;
        for (int64_t b = 0; b < B; ++b)
            for (int p = 0; p < HW; ++p) {
                const float *e = E[c] + idx[(b * C + c) * HW + p] * d;
                for (int j = 0; j < d; ++j) {
                    float zv = z[(b * Dtot + (long)c * cs + j) * HW + p];
                    float diff = e[j] - zv;
                    q_out[(b * (long)C * d + (long)c * d + j) * HW + p] = zv + diff;
                    acc += (double)(diff * diff);
                }
            }
        float m = (float)(acc / ((double)B * HW * d));
        loss_out[c] = m * beta + m;
        total = total + loss_out[c];
    }
    loss_out[C] = total;
}

/* autograd of models/vq_vae.py:43-53; overlap accumulation for models/mcq_vae.py:117 slices */
void ctvq_c_backward(const float *z, const float *const *E, const int64_t *idx, const float *g_out, float g_loss,
                     int64_t B, int Dtot, int HW, int C, int d, int K, int cs, float beta, float *gz /* [B,Dtot,HW] */,
                     float *gE /* [C,K,d] */) {
    double nd = (double)B * HW * d;
    memset(gz, 0, sizeof(float) * (size_t)(B * Dtot * HW));
    double *acc = (double *)calloc((size_t)C * K * d, sizeof(double));
    for (int c = 0; c < C; ++c)
        for (int64_t b = 0; b < B; ++b)
            for (int p = 0; p < HW; ++p) {
                int64_t k = idx[(b * C + c) * HW + p];
                const float *e = E[c] + k * d;
                for (int j = 0; j < d; ++j) {
                    long zi = (b * Dtot + (long)c * cs + j) * HW + p;
                    float diff = e[j] - z[zi];
                    gz[zi] += g_out[(b * (long)C * d + (long)c * d + j) * HW + p] +
                              (float)(-2.0 * beta / nd) * g_loss * diff;
                    acc[((long)c * K + k) * d + j] += (double)diff;
                }
            }
    for (long i = 0; i < (long)C * K * d; ++i) gE[i] = (float)((2.0 / nd) * g_loss * acc[i]);
    free(acc);
}

/* models/vanilla_vae.py:115-117 (eps supplied) and :143 */
void ctvq_c_reparam_kld(const float *mu, const float *lv, const float *eps, int64_t B, int L, float *z_out,
                        float *kld_out) {
    double acc = 0.0;
    for (int64_t i = 0; i < B * L; ++i) {
        z_out[i] = eps[i] * expf(0.5f * lv[i]) + mu[i];
        acc += (double)(1.0f + lv[i] - mu[i] * mu[i] - expf(lv[i]));
    }
    *kld_out = (float)(-0.5 * acc / (double)B);
}

/* ---- CT-mode codec, same layouts as the kernels: one-hots / scores fp32 [B, K, S] contiguous
 * (S = C*H*W), indices int64 [B, S] ------------------------------------------------------------------------------ */

/* CTMCQVAE.ct_preprocess, models/ct_mcq_vae.py:472-483 (F.one_hot + view + permute, made contiguous) */
void ctvq_c_onehot(const int64_t *idx, int64_t B, int64_t S, int K, float *out) {
    for (int64_t b = 0; b < B; ++b)
        for (int k = 0; k < K; ++k)
            for (int64_t s = 0; s < S; ++s) out[(b * K + k) * S + s] = idx[b * S + s] == k ? 1.0f : 0.0f;
}

/* CTMCQVAE.ct_postprocess, models/ct_mcq_vae.py:485-496: torch.argmax over the class dimension
 * (first maximum wins; the first NaN wins over any number) */
void ctvq_c_class_argmax(const float *x, int64_t B, int64_t S, int K, int64_t *idx) {
    for (int64_t b = 0; b < B; ++b)
        for (int64_t s = 0; s < S; ++s) {
            float best = x[(b * K) * S + s];
            int bi = 0;
            for (int k = 1; k < K; ++k) {
                const float v = x[(b * K + k) * S + s];
                if (v > best || (v != v && best == best)) { best = v; bi = k; }
            }
            idx[b * S + s] = bi;
        }
}

/* CausalTransition.latent_CrossEntropy_loss, models/ct_mcq_vae.py:306-311, and its gradient w.r.t. latent:
 *   loss = mean_rows( log(sum_k x'_k) - log(x'_t) ),  x' = max(x, 1e-4),  t = argmax_k y
 *   g_x[b,k,s] = g / R * [x >= 1e-4] * (1/sum' - [k == t]/x'_t)
 * Sums in double (the checker wants the exact value; the kernels are held to 1e-5 relative). */
void ctvq_c_latent_ce(const float *x, const float *y, int64_t B, int64_t S, int K, float g, float *loss_out, float *gx) {
    const double R = (double)B * (double)S;
    double total = 0.0;
    for (int64_t b = 0; b < B; ++b)
        for (int64_t s = 0; s < S; ++s) {
            float best = y[(b * K) * S + s];
            int t = 0;
            double sum = 0.0;
            for (int k = 0; k < K; ++k) {
                const float xv = x[(b * K + k) * S + s];
                sum += (double)(xv != xv ? xv : (xv > 1e-4f ? xv : 1e-4f));
                const float yv = y[(b * K + k) * S + s];
                if (k > 0 && (yv > best || (yv != yv && best == best))) { best = yv; t = k; }
            }
            const float xt_raw = x[(b * K + t) * S + s];
            const double xt = (double)(xt_raw != xt_raw ? xt_raw : (xt_raw > 1e-4f ? xt_raw : 1e-4f));
            total += log(sum) - log(xt);
            if (gx)
                for (int k = 0; k < K; ++k) {
                    const float xv = x[(b * K + k) * S + s];
                    double d = 1.0 / sum;
                    if (k == t) d -= 1.0 / xt;
                    gx[(b * K + k) * S + s] = xv >= 1e-4f ? (float)((double)g / R * d) : 0.0f;
                }
        }
    *loss_out = (float)(total / R);
}
