"""Round-2 golden vectors, minted by running the UNMODIFIED reference imported live from /root/reference
(``python tests/golden/make_golden_r2.py`` in the build container; the tree does not exist on the GPU box).

What they add to make_golden.py's set:
  * exact ties and latents-equal-to-codewords in shapes the tcgen05 kernels COVER (the round-1 edge goldens were all
    too small for them): the config-2 multi-codebook shape (vq_fwd_tc_fast_kernel), K=512 x D=64 at HW=256 (streaming
    kernel), the config-3 single-codebook shape (resident-codebook kernel) and a two-codebook shape that takes the generic
    tcgen05 kernel.  Reference entry points: MultipleCodebookVectorQuantizer.forward (models/mcq_vae.py:130-137) and
    VectorQuantizerMS.forward (:67-74); the first-minimum rule under test is torch.argmin's (:37).
  * non-finite latents (NaN, +-inf, |z|^2 overflow, an all-+inf distance row) in the same shapes: indices only
    (models/mcq_vae.py:26-39, torch.argmin: the first NaN wins, an all-inf row answers 0).
  * BetaVAE.loss_function (models/beta_vae.py:130-152), loss types 'H' and 'B', over successive calls so the capacity
    schedule C = clamp(C_max / C_stop_iter * num_iter, 0, C_max) is pinned, and VanillaVAE.loss_function
    (models/vanilla_vae.py:128-146) at a second shape.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_live  # noqa: E402

torch.set_num_threads(1)  # fixed reduction order for reproducible goldens
models = ref_live.load()
from models.mcq_vae import MultipleCodebookVectorQuantizer, VectorQuantizerMS  # noqa: E402


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


def dup_codebook(K, d, scale=0.5):
    """Second half of the rows duplicates the first half: every row has an exact tie, the FIRST copy must win."""
    half = torch.randn(K // 2, d) * scale
    return torch.cat([half, half], 0)


def plant_codewords(z, books, d, every=3):
    """Overwrite every `every`-th latent vector's slice for ONE codebook (cycling) with a codeword of that codebook, so
    its distance is ~0 (heavy cancellation) and ties with the duplicate row."""
    B, D, H, W = z.shape
    C = len(books)
    n = 0
    with torch.no_grad():
        for b in range(B):
            for p in range(0, H * W, every):
                c = n % C
                k = (7 * n + 3) % books[c].shape[0]
                z[b, c:c + d, p // W, p % W] = books[c][k]
                n += 1
    return z


def run_tie_mcq(name, seed, K, D, C, shape, beta=0.25, g_loss=0.7):
    torch.manual_seed(seed)
    d = D // C
    m = MultipleCodebookVectorQuantizer(K, D, C, beta)
    for q in m.quantizers:
        q.embedding.weight.data = dup_codebook(K, d)
    z = plant_codewords(torch.randn(*shape), [q.embedding.weight.data for q in m.quantizers], d).requires_grad_(True)
    out, loss, inds = m(z, inds=True)
    g_out = torch.randn_like(out)
    ((out * g_out).sum() + g_loss * loss).backward()
    arrs = dict(z=z, beta=beta, C=C, inds=inds, out=out, loss=loss, g_out=g_out, g_loss=g_loss, gz=z.grad)
    for i, q in enumerate(m.quantizers):
        arrs[f"codebook{i}"] = q.embedding.weight
        arrs[f"gE{i}"] = q.embedding.weight.grad
        with torch.no_grad():
            arrs[f"loss{i}"] = q.compute_latents(z[:, i:i + d], inds[:, i])[1]
    print(f"  {name}: reference picked a second-half (duplicate) row on {int((inds >= K // 2).sum())} of {inds.numel()} rows")
    save(name, **arrs)


def run_tie_single(name, seed, K, D, shape, beta=0.25, g_loss=0.7):
    torch.manual_seed(seed)
    m = VectorQuantizerMS(K, D, beta)
    m.embedding.weight.data = dup_codebook(K, D)
    z = plant_codewords(torch.randn(*shape), [m.embedding.weight.data], D).requires_grad_(True)
    out, loss, inds = m(z, inds=True)
    g_out = torch.randn_like(out)
    ((out * g_out).sum() + g_loss * loss).backward()
    print(f"  {name}: reference picked a second-half (duplicate) row on {int((inds >= K // 2).sum())} of {inds.numel()} rows")
    save(name, z=z, codebook=m.embedding.weight, beta=beta, inds=inds, out=out, loss=loss, g_out=g_out, g_loss=g_loss,
         gz=z.grad, gE=m.embedding.weight.grad)


def poison(z, d):
    """Rows with non-finite distances; the rest of the tensor stays ordinary."""
    z[0, 3, 1, 2] = float("nan")
    z[1, 0, 0, 0] = float("inf")
    z[2, d - 1, 7, 7] = float("-inf")
    z[3, 5, 4, 4] = 3.0e38          # |z|^2 overflows to +inf, z.e stays finite: every distance is +inf -> index 0
    z[3, 6, 4, 4] = -3.0e38
    z[4, 2, 0, 5] = float("nan")
    z[4, 2, 0, 6] = float("inf")
    return z


def run_nonfinite_mcq(name, seed, K, D, C, shape):
    torch.manual_seed(seed)
    d = D // C
    m = MultipleCodebookVectorQuantizer(K, D, C, 0.25)
    for q in m.quantizers:
        q.embedding.weight.data = torch.randn(K, d) * 0.5
    z = poison(torch.randn(*shape), d)
    with torch.no_grad():
        inds = m.compute_inds(z)
    arrs = dict(z=z, C=C, inds=inds)
    for i, q in enumerate(m.quantizers):
        arrs[f"codebook{i}"] = q.embedding.weight
    save(name, **arrs)


def run_nonfinite_single(name, seed, K, D, shape):
    torch.manual_seed(seed)
    m = VectorQuantizerMS(K, D, 0.25)
    m.embedding.weight.data = torch.randn(K, D) * 0.5
    z = poison(torch.randn(*shape), D)
    with torch.no_grad():
        inds = m.compute_inds(z)
    save(name, z=z, codebook=m.embedding.weight, inds=inds)


# ---- exact ties / latents on codewords in tcgen05-covered shapes -----------------------------------------------------
run_tie_mcq("tie_tc_mcq_cfg2", 21, 64, 128, 4, (8, 128, 8, 8))            # vq_fwd_tc_fast_kernel (C=4, d=32, K=64, HW=64)
run_tie_single("tie_tc_stream_k512", 22, 512, 64, (2, 64, 16, 16))        # streaming kernel (K=512, D=64, HW=256)
run_tie_single("tie_tc_c1_cfg3", 23, 64, 128, (8, 128, 8, 8), beta=0.1)   # resident single-codebook kernel (config 3)
run_tie_mcq("tie_tc_generic_c2", 24, 64, 64, 2, (4, 64, 8, 8))            # generic tcgen05 kernel (C=2, d=32)
run_tie_single("tie_tc_stream_d32_k256", 25, 256, 32, (2, 32, 16, 16))    # streaming kernel, 4 teams
# ---- non-finite rows in the same shapes (names start with "nonfinite_": indices only) --------------------------------
run_nonfinite_mcq("nonfinite_mcq_cfg2", 31, 64, 128, 4, (8, 128, 8, 8))
run_nonfinite_single("nonfinite_stream_k512", 32, 512, 64, (5, 64, 16, 16))
run_nonfinite_single("nonfinite_c1_cfg3", 33, 64, 128, (8, 128, 8, 8))

# ---- Gaussian losses: VanillaVAE / BetaVAE loss_function (models/vanilla_vae.py:128-146, models/beta_vae.py:130-152) -
torch.manual_seed(41)
B, L = 32, 128
mu, lv = torch.randn(B, L), torch.randn(B, L) * 0.5
rec, inp = torch.randn(B, 3, 8, 8), torch.randn(B, 3, 8, 8)
vv = models.VanillaVAE(3, L)
r = vv.loss_function(rec, inp, mu, lv, M_N=0.00025)
arrs = dict(mu=mu, logvar=lv, recons=rec, input=inp, M_N=0.00025, vanilla_loss=r["loss"],
            vanilla_recons=r["Reconstruction_Loss"], vanilla_KLD=r["KLD"])
for lt in ("H", "B"):
    models.BetaVAE.num_iter = 0
    bv = models.BetaVAE(3, L, beta=4, gamma=10.0, max_capacity=25, Capacity_max_iter=3, loss_type=lt)
    for it in range(1, 6):  # the 'B' capacity C = 25/3 * num_iter saturates at 25 from the third call on
        r = bv.loss_function(rec, inp, mu, lv, M_N=0.00025)
        assert bv.num_iter == it
        arrs[f"beta{lt}_loss_{it}"] = r["loss"]
        arrs[f"beta{lt}_recons_{it}"] = r["Reconstruction_Loss"]
        arrs[f"beta{lt}_KLD_{it}"] = r["KLD"]
arrs.update(beta=4, gamma=10.0, max_capacity=25, Capacity_max_iter=3)
save("gaussian_losses", **arrs)
