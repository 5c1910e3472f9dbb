/* ctvq — C ABI of the B200-native codebook quantiser (drop-in boundary).
 *
 * The reference (Strong-AI-Lab/ct-vae) has no FFI of its own: the boundary it exposes is the Python
 * nn.Module surface of models/vq_vae.py:7-55 and models/mcq_vae.py:7-137.  Those modules are mirrored in
 * ct_vae_b200/modules.py, which binds THIS header through ctypes.  Every entry point below names the
 * reference code it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.  All data pointers are DEVICE pointers to
 *     caller-owned memory on `device`; the compute entry points allocate nothing and are re-entrant (the caller
 *     passes a zero-initialised workspace per stream, ctvq_workspace_bytes()).  PROCESS-GLOBAL state, all of it
 *     set-once or test-only: the kernel-path override (ctvq_set_path: one atomic int, for tests and A/B runs), the
 *     dlopen()ed NCCL function table (ctvq_nccl_load), the cached SM count per device, function attributes
 *     (cudaFuncSetAttribute), the A/B environment switches CTVQ_BWD_NO_TMA / CTVQ_PEER_TIMEOUT_MS (read once), and
 *     the two debug hooks at the end of this header.  ctvq_last_path() is per host thread.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised (CUDA-graph safe).
 *   - return: 0 = OK, <0 = bad argument / unsupported shape (CTVQ_E_*), >0 = a cudaError_t.
 *     ctvq_strerror() renders either.
 *   - layout: latents NCHW = [B, Dtot, HW] contiguous; codebook c is [K, d] row-major and reads input
 *     channels c*chan_stride .. c*chan_stride+d-1 (the reference's `latents[:, i:i+d]`,
 *     models/mcq_vae.py:104,117, is chan_stride = 1: overlapping slices); quantised output
 *     [B, C*d, HW]; indices int64 [B, C, HW].
 *   - dtype: CTVQ_F32 (the reference's only arithmetic type).  Distances are evaluated as
 *     fl(fl(|z|^2 + |e_k|^2) - 2*dot) with every sum a sequential FMA chain over ascending channel, and
 *     argmin takes the first minimum (NaN: first NaN) — see DESIGN.md "arithmetic contract".
 */
#ifndef CTVQ_H_
#define CTVQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTVQ_VERSION 200          /* 0.2.0: near-tie counter, error-flag read-back, bf16 */
#define CTVQ_MAX_CODEBOOKS 64
#define CTVQ_MAX_SEGMENTS 4

#define CTVQ_F32 0
#define CTVQ_BF16 1

#define CTVQ_OK 0
#define CTVQ_E_BADARG (-1)        /* null pointer, non-positive size, C*... inconsistent */
#define CTVQ_E_UNSUPPORTED (-2)   /* shape or dtype this build has no kernel for */
#define CTVQ_E_WORKSPACE (-3)     /* workspace too small */
#define CTVQ_E_NCCL (-4)          /* NCCL not loaded / NCCL call failed */
#define CTVQ_E_NOT_BUILT (-5)

/* kernel selection for ctvq_argmin / ctvq_forward */
#define CTVQ_PATH_AUTO 0
#define CTVQ_PATH_SIMT 1          /* shared-memory-staged fp32 FFMA kernel */
#define CTVQ_PATH_TC 2            /* tcgen05/TMEM distance GEMM + exact fp32 re-scoring */
#define CTVQ_PATH_TC_STREAM 3     /* force the streaming single-codebook tcgen05 kernel (tests / A-B; reported as CTVQ_PATH_TC) */

int ctvq_version(void);
const char* ctvq_strerror(int rc);

/* Bytes of zero-initialised device workspace one stream needs: a 1 KB self-cleaning header, plus -- for C == 1 -- the
 * scratch of the streaming single-codebook kernel (|e_k|^2 and its tf32-split GEMM blocks: K*4 + ceil(K/256)*8 KB,
 * rewritten each call).  Passing only ctvq_workspace_bytes(0,0,0) is valid: such calls take the other kernels. */
size_t ctvq_workspace_bytes(int C, int K, int d);

/* Index validation.  Every entry point that CONSUMES caller-supplied indices (ctvq_gather_st_loss, ctvq_backward,
 * ctvq_onehot_from_inds) clamps an index outside [0, K) and sets bit 0 of an error word in the workspace header instead
 * of faulting (the reference raises from scatter_ / F.one_hot, models/vq_vae.py:40, models/ct_mcq_vae.py:480).  This
 * call copies the word to the host and clears it; it SYNCHRONISES `stream`.  *err_out != 0 -> the caller raises. */
int ctvq_read_and_clear_err(void* workspace, size_t ws_bytes, unsigned* err_out_host, int device, void* stream);

/* Force a kernel path for subsequent calls of the WHOLE PROCESS (one atomic int; tests / A-B runs only -- production code
 * leaves it at AUTO, which picks by shape).  Returns the previous value. */
int ctvq_set_path(int path);
/* Which path the last ctvq_argmin/ctvq_forward call on this host thread dispatched to. */
int ctvq_last_path(void);

/* NEAR-TIE ACCOUNTING (BASELINE.json north_star: "near-ties with a relative top-2 distance gap below 1e-6 are counted
 * and reported, not hidden"; the distances are those of models/vq_vae.py:30-32, the winner that of :35).
 * `neartie_count_out` (may be NULL) points to ONE device counter that ctvq_argmin / ctvq_forward ADD to: the number of
 * (row, codebook) pairs whose best and second-best fp32 distances d1 <= d2 (arithmetic contract above) satisfy
 * d2 - d1 <= 1e-6 * |d1| -- exact ties included, also at distance 0.  These are exactly the rows where a different summation order (the
 * reference's sgemm) may legitimately pick the other code.  The caller zero-initialises the counter and may let it
 * accumulate over many calls (no synchronisation here).  Rows with a non-finite distance are not counted. */
#define CTVQ_NEAR_TIE_REL 1e-6f

/* compute_inds — replaces VectorQuantizerMS.compute_inds (models/mcq_vae.py:26-39),
 * MultipleCodebookVectorQuantizer.compute_inds (:100-110) and the distance+argmin half of
 * VectorQuantizer.forward (models/vq_vae.py:25-35).  `n_seg` input tensors of identical shape are
 * processed in ONE launch (the x / y pair of CTMCQVAE.forward_action/_causal,
 * models/ct_mcq_vae.py:530,536,555-556); pass n_seg = 1 otherwise. */
int ctvq_argmin(const void* const* z_segs, int n_seg, const void* const* codebooks, int64_t B, int Dtot,
                int HW, int C, int d, int K, int chan_stride, int dtype, int64_t* const* idx_out_segs,
                unsigned long long* neartie_count_out, void* workspace, size_t ws_bytes, int device, void* stream);

/* compute_latents — replaces VectorQuantizerMS.compute_latents (models/mcq_vae.py:41-64) and
 * MultipleCodebookVectorQuantizer.compute_latents (:112-127): gather by CALLER-SUPPLIED indices,
 * loss_c = m*beta + m with m = mean((q - z)^2), straight-through output z + (q - z).
 * loss_out has C+1 floats: per-codebook losses then their left-to-right sum (mcq_vae.py:125). */
int ctvq_gather_st_loss(const void* z, const void* const* codebooks, const int64_t* idx, int64_t B,
                        int Dtot, int HW, int C, int d, int K, int chan_stride, int dtype, float beta,
                        void* q_out, float* loss_out, void* workspace, size_t ws_bytes, int device,
                        void* stream);

/* forward — replaces VectorQuantizer.forward (models/vq_vae.py:24-55), VectorQuantizerMS.forward
 * (models/mcq_vae.py:67-74) and MultipleCodebookVectorQuantizer.forward (:130-137): argmin + gather +
 * loss + straight-through in one pass over z (the N x K distance matrix never reaches HBM). */
int ctvq_forward(const void* z, const void* const* codebooks, int64_t B, int Dtot, int HW, int C, int d,
                 int K, int chan_stride, int dtype, float beta, int64_t* idx_out, void* q_out,
                 float* loss_out, unsigned long long* neartie_count_out, void* workspace, size_t ws_bytes,
                 int device, void* stream);

/* backward — replaces autograd through models/vq_vae.py:43-53:
 *   gz[b,ch,p] = sum over (c,j) with c*chan_stride+j == ch of
 *                g_out[b,c*d+j,p] + g_loss*beta*2*(z - q)/(N*d)          (zero where no slice reads ch)
 *   gE[c,k,:]  = g_loss * 2/(N*d) * sum_{n: idx=k} (q_n - z_n)
 * g_loss points to ONE device float (the gradient of the summed vq_loss); g_out may be NULL (treated as
 * zero, e.g. the loss-only path).  gE_out [C,K,d] is overwritten (zeroed then accumulated). */
int ctvq_backward(const void* z, const void* const* codebooks, const int64_t* idx, const void* g_out,
                  const float* g_loss, int64_t B, int Dtot, int HW, int C, int d, int K, int chan_stride,
                  int dtype, float beta, void* gz_out, float* gE_out, void* workspace, size_t ws_bytes,
                  int device, void* stream);

/* Gaussian branch — replaces VanillaVAE/BetaVAE.reparameterize (models/vanilla_vae.py:107-117,
 * models/beta_vae.py:112-122) with eps supplied by the caller, fused with the KL term of
 * loss_function (vanilla_vae.py:143, beta_vae.py:141): z = eps*exp(0.5*lv) + mu,
 * kld = mean_b(-0.5 * sum_l(1 + lv - mu^2 - exp(lv))).  mu/logvar/eps/z are [B, L] fp32. */
int ctvq_reparam_kld_fwd(const float* mu, const float* logvar, const float* eps, int64_t B, int L,
                         float* z_out, float* kld_out, void* workspace, size_t ws_bytes, int device,
                         void* stream);
/* g_z [B,L] may be NULL; g_kld points to one device float (may be NULL = 0). */
int ctvq_reparam_kld_bwd(const float* mu, const float* logvar, const float* eps, const float* g_z,
                         const float* g_kld, int64_t B, int L, float* g_mu_out, float* g_logvar_out,
                         int device, void* stream);

/* CT-mode codec -- the converters either side of the quantiser in CausalTransition mode.
 * One-hots are fp32 [B, K, S] contiguous with S = C*H*W (the reference's [B, N, K*H, W] tensor after its permute,
 * models/ct_mcq_vae.py:481-482), indices int64 [B, S] (= [B, C, H, W]).
 *   ctvq_onehot_from_inds  replaces CTMCQVAE.ct_preprocess  (models/ct_mcq_vae.py:472-483): F.one_hot + view + permute.
 *                          An index outside [0, K) leaves its row all-zero and is flagged in the workspace (the
 *                          reference raises from F.one_hot).
 *   ctvq_inds_from_onehot  replaces CTMCQVAE.ct_postprocess (models/ct_mcq_vae.py:485-496): permute + reshape +
 *                          torch.argmax over the class dimension (first maximum wins, the first NaN wins).
 *   ctvq_latent_ce_fwd/bwd replace CausalTransition.latent_CrossEntropy_loss (models/ct_mcq_vae.py:306-311):
 *                          loss = mean_rows( log(sum_k x'_k) - log(x'_t) ), x' = max(latent, 1e-4), t = argmax_k latent_y;
 *                          fwd also returns the targets [B,S] and the row sums [B,S] the backward re-uses;
 *                          bwd: g_latent = g_loss/(B*S) * [latent >= 1e-4] * (1/sum' - [k==t]/x'_t); latent_y gets no
 *                          gradient (it is detached at models/ct_mcq_vae.py:300). */
int ctvq_onehot_from_inds(const int64_t* idx, int64_t B, int64_t S, int K, float* onehot_out, void* workspace,
                          size_t ws_bytes, int device, void* stream);
int ctvq_inds_from_onehot(const float* scores, int64_t B, int64_t S, int K, int64_t* idx_out, int device, void* stream);
int ctvq_latent_ce_fwd(const float* latent, const float* latent_y, int64_t B, int64_t S, int K, int64_t* target_out,
                       float* rowsum_out, float* loss_out, void* workspace, size_t ws_bytes, int device, void* stream);
int ctvq_latent_ce_bwd(const float* latent, const int64_t* target, const float* rowsum, const float* g_loss, int64_t B,
                       int64_t S, int K, float* g_latent_out, int device, void* stream);

/* Codebook-gradient all-reduce — the one collective of the path (replaces the share of Lightning's DDP
 * bucket all-reduce that carries vq_layer.*.embedding.weight.grad, run.py:99).  NCCL is dlopen()ed from
 * `libnccl_path` (the torch-bundled libnccl.so.2); communicators are created from a 128-byte unique id
 * the caller broadcasts by its own means. */
int ctvq_nccl_load(const char* libnccl_path);
int ctvq_nccl_unique_id(void* id128_out);
int ctvq_nccl_comm_init(void** comm_out, int nranks, int rank, const void* id128, int device);
int ctvq_nccl_comm_destroy(void* comm);
/* sum over ranks then multiply by `scale` (1/world = DDP's gradient averaging), in place, on `stream`. */
int ctvq_allreduce_codebook_grad(void* comm, float* gE, size_t count, float scale, int device, void* stream);

/* One-shot all-reduce of the codebook gradient over NVLink peer memory (no NCCL on the path), normally FUSED into the
 * backward kernel: every rank owns a "symmetric" buffer (2 parities x `world` receive slots of `count_max` floats +
 * `world` flag words) that its peers map through CUDA IPC.  Push protocol (csrc/ctvq_peer.cuh): once the local partial
 * gradient is complete, ONE CTA (1) stores it into slot [epoch parity][rank] of every rank's buffer (posted NVLink
 * writes), (2) fences and posts `epoch` into every rank's flag row, (3) waits until its own flag row shows `epoch` for
 * every rank (local polling), (4) sums the `world` local slots in rank order, times `scale`, into `out` -- bit-identical
 * on every rank.  Slots alternate by epoch parity, which makes a second barrier unnecessary; epochs must increase by one
 * per collective on every rank, starting at 1.  A peer that does not arrive within CTVQ_PEER_TIMEOUT_MS (env, default
 * 30000) does not trap: bit 1 of the workspace error word is set (ctvq_read_and_clear_err) and the kernel finishes.
 * The handle exchange is the caller's job (64-byte cudaIpcMemHandle_t per rank); count_max must be a multiple of 4. */
#define CTVQ_MAX_PEERS 8
#define CTVQ_IPC_HANDLE_BYTES 64
size_t ctvq_peer_buffer_bytes(size_t count_max, int world);
int ctvq_peer_alloc(void** dev_ptr_out, size_t count_max, int world, int device);
int ctvq_peer_free(void* dev_ptr, int device);
int ctvq_peer_export(void* dev_ptr, void* handle64_out, int device);
int ctvq_peer_import(const void* handle64, void** dev_ptr_out, int device);
int ctvq_peer_close(void* dev_ptr, int device);
/* Stand-alone collective on an already complete local gradient `src` [count] (one 512-thread CTA).
 * peer_bufs: host array of `world` device pointers (entry `rank` = this rank's own buffer). */
int ctvq_peer_allreduce(void* const* peer_bufs, int world, int rank, size_t count_max, const float* src, size_t count,
                        unsigned epoch, float scale, float* out, void* workspace, size_t ws_bytes, int device,
                        void* stream);
/* backward + all-reduce in ONE launch: ctvq_backward (above) whose last CTA runs the collective on the finished
 * gE_local [C,K,d] and leaves sum_over_ranks(gE) * scale in gE_reduced_out [C,K,d] (scale = 1/world reproduces DDP's
 * gradient averaging, run.py:99).  Works with whichever backward kernel the shape dispatches to. */
int ctvq_backward_allreduce(const void* z, const void* const* codebooks, const int64_t* idx, const void* g_out,
                            const float* g_loss, int64_t B, int Dtot, int HW, int C, int d, int K, int chan_stride,
                            int dtype, float beta, void* gz_out, float* gE_local, void* const* peer_bufs, int world,
                            int rank, size_t count_max, unsigned epoch, float scale, float* gE_reduced_out,
                            void* workspace, size_t ws_bytes, int device, void* stream);

/* Debug hooks (development aids, not part of the product surface; both are process-global pointers, null = off):
 *   ctvq_debug_set_fast_trace  8 %globaltimer stamps per CTA of vq_fwd_tc_fast_kernel into buf[SMs*8] (tools/trace_fast.py)
 *   ctvq_debug_set_tc_dump     raw TMEM dot products of the generic tcgen05 kernel into buf (tools/debug_tc.py) */
void ctvq_debug_set_fast_trace(unsigned long long* buf);
void ctvq_debug_set_tc_dump(float* buf);

#ifdef __cplusplus
}
#endif
#endif /* CTVQ_H_ */
