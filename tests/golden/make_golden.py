"""Mint the golden vectors by running the UNMODIFIED reference, imported live from /root/reference.

Run once in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Each ``*.npz`` holds the exact inputs and the reference's outputs (indices, straight-through output,
loss, and autograd gradients for an upstream gradient ``g_out`` on the output and weight ``g_loss`` on
the loss).  Seeds follow the reference configs: 1265 (configs/vq_vae.yaml:22), 1320
(configs/mcq_vae.yaml:27), 1250 (configs/ct_mcq_vae.yaml:35).  Reference entry points exercised:
``VectorQuantizer.forward`` (models/vq_vae.py:24-55), ``VectorQuantizerMS.compute_inds /
compute_latents / forward`` (models/mcq_vae.py:26-74), ``MultipleCodebookVectorQuantizer.*``
(models/mcq_vae.py:100-137), ``VQVAE.forward/loss_function`` (models/vq_vae.py:189-211, recipe of
tests/test_vq_vae.py:17-29), ``VanillaVAE.reparameterize`` math and KLD (models/vanilla_vae.py:115-117,143).
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
from models.vq_vae import VectorQuantizer  # noqa: E402
from models.mcq_vae import VectorQuantizerMS, MultipleCodebookVectorQuantizer  # noqa: E402


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


def run_single(name, seed, K, D, shape, beta=0.25, codebook="init", cls=VectorQuantizer, z_fn=None, g_loss=0.7):
    torch.manual_seed(seed)
    m = cls(K, D, beta)
    if codebook == "trained":
        m.embedding.weight.data = torch.randn(K, D) * 0.5
    elif isinstance(codebook, torch.Tensor):
        m.embedding.weight.data = codebook.clone()
    z = (z_fn() if z_fn else torch.randn(*shape)).requires_grad_(True)
    out, loss = m(z)
    g_out = torch.randn_like(out)
    ((out * g_out).sum() + g_loss * loss).backward()
    with torch.no_grad():
        flat = z.permute(0, 2, 3, 1).reshape(-1, D)
        w = m.embedding.weight
        dist = torch.sum(flat ** 2, 1, keepdim=True) + torch.sum(w ** 2, 1) - 2 * flat @ w.t()
        inds = torch.argmin(dist, 1).view(z.shape[0], z.shape[2], z.shape[3])
    save(name, z=z, codebook=m.embedding.weight, beta=beta, inds=inds, out=out, loss=loss,
         g_out=g_out, g_loss=g_loss, gz=z.grad, gE=m.embedding.weight.grad)


def run_mcq(name, seed, K, D, C, shape, beta=0.25, codebook="init", g_loss=0.7, external_inds=False):
    torch.manual_seed(seed)
    m = MultipleCodebookVectorQuantizer(K, D, C, beta)
    d = D // C
    if codebook == "trained":
        for q in m.quantizers:
            q.embedding.weight.data = torch.randn(K, d) * 0.5
    z = torch.randn(*shape).requires_grad_(True)
    if external_inds:  # CT 'base' mode: indices come from the transition layer (ct_mcq_vae.py:519)
        inds = torch.randint(0, K, (shape[0], C, shape[2], shape[3]))
        out, loss = m.compute_latents(z, inds)
    else:
        out, loss, inds = m(z, inds=True)
        assert torch.equal(inds, m.compute_inds(z))
    g_out = torch.randn_like(out)
    ((out * g_out).sum() + g_loss * loss).backward()
    arrs = dict(z=z, beta=beta, C=C, inds=inds, out=out, loss=loss, g_out=g_out, g_loss=g_loss, gz=z.grad)
    for i, q in enumerate(m.quantizers):
        arrs[f"codebook{i}"] = q.embedding.weight
        arrs[f"gE{i}"] = q.embedding.weight.grad
        with torch.no_grad():
            arrs[f"loss{i}"] = q.compute_latents(z[:, i:i + d], inds[:, i])[1]
    save(name, **arrs)


# ---- config 1: VQ-VAE quantiser, K=512 D=64, latents [B,64,16,16] (configs/vq_vae.yaml) -------------
run_single("vq_cfg1_init", 1265, 512, 64, (4, 64, 16, 16))
run_single("vq_cfg1_trained", 1265, 512, 64, (4, 64, 16, 16), codebook="trained")
# ---- MS variant, same maths through compute_inds/compute_latents (mcq_vae.py:67-74) -----------------
run_single("vqms_small", 7, 32, 16, (3, 16, 4, 4), cls=VectorQuantizerMS, codebook="trained")
# ---- config 2: MCQ-VAE quantiser C=4 d=32 K=64, latents [B,128,8,8] (configs/mcq_vae.yaml) ---------
run_mcq("mcq_cfg2_init", 1320, 64, 128, 4, (8, 128, 8, 8))
run_mcq("mcq_cfg2_trained", 1320, 64, 128, 4, (8, 128, 8, 8), codebook="trained")
# ---- config 3: CT-MCQ-VAE quantiser C=1 d=128 K=64 beta=0.1, [16,128,8,8] (configs/ct_mcq_vae.yaml) -
run_mcq("ct_cfg3_trained", 1250, 64, 128, 1, (16, 128, 8, 8), beta=0.1, codebook="trained")
run_mcq("ct_cfg3_external_inds", 1250, 64, 128, 1, (16, 128, 8, 8), beta=0.1, codebook="trained", external_inds=True)
run_mcq("mcq_external_inds", 11, 16, 24, 3, (2, 24, 3, 5), codebook="trained", external_inds=True)
# ---- edge cases -------------------------------------------------------------------------------------
# exact ties: duplicated codebook rows -> the FIRST index must win (torch.argmin rule)
torch.manual_seed(3)
dup = torch.randn(8, 6)
dup = torch.cat([dup, dup[:4], dup], 0)  # rows 8..11 duplicate 0..3, rows 12..19 duplicate 0..7
run_single("edge_ties", 3, 20, 6, (2, 6, 3, 3), codebook=dup)
# latents that ARE codewords (distance ~0, heavy cancellation)
run_single("edge_on_codeword", 5, 20, 6, (2, 6, 2, 5), codebook=dup,
           z_fn=lambda: dup[torch.randint(0, 20, (2, 2, 5))].permute(0, 3, 1, 2).contiguous())
# ragged shapes: HW=1, odd D, K=1, HW not a multiple of 4
run_single("edge_hw1", 9, 3, 8, (5, 8, 1, 1), codebook="trained")
run_single("edge_k1", 9, 1, 5, (2, 5, 3, 5), codebook="trained")
run_single("edge_odd", 9, 37, 5, (3, 5, 3, 5), codebook="trained")
run_mcq("edge_mcq_odd", 13, 7, 15, 5, (2, 15, 3, 3), codebook="trained")

# ---- the reference's own test recipe: VQVAE(3,64,512), randn(16,3,64,64), M_N=0.005 ----------------
torch.manual_seed(1265)
vae = models.VQVAE(3, 64, 512)
x = torch.randn(16, 3, 64, 64)
with torch.no_grad():
    enc = vae.encode(x)[0]
    q, vq_loss = vae.vq_layer(enc)
    res = vae.loss_function(vae.decode(q), x, vq_loss, M_N=0.005)
    flat = enc.permute(0, 2, 3, 1).reshape(-1, 64)
    w = vae.vq_layer.embedding.weight
    dist = torch.sum(flat ** 2, 1, keepdim=True) + torch.sum(w ** 2, 1) - 2 * flat @ w.t()
    inds = torch.argmin(dist, 1).view(16, 16, 16)
save("vqvae_recipe_encoder_latents", z=enc, codebook=w, beta=0.25, inds=inds, out=q, loss=vq_loss,
     model_loss=res["loss"], recons_loss=res["Reconstruction_Loss"])

# ---- Gaussian branch (config 5 shape family, small) -----------------------------------------------
torch.manual_seed(1265)
mu = torch.randn(64, 128, requires_grad=True)
lv = (torch.randn(64, 128) * 0.5).requires_grad_(True)
eps = torch.randn(64, 128)
std = torch.exp(0.5 * lv)
zs = eps * std + mu  # vanilla_vae.py:115-117 with eps supplied
kld = torch.mean(-0.5 * torch.sum(1 + lv - mu ** 2 - lv.exp(), dim=1), dim=0)  # vanilla_vae.py:143
g_z = torch.randn_like(zs)
g_kld = 0.00025
((zs * g_z).sum() + g_kld * kld).backward()
save("reparam_kld", mu=mu, logvar=lv, eps=eps, z=zs, kld=kld, g_z=g_z, g_kld=g_kld, g_mu=mu.grad, g_logvar=lv.grad)
# the real modules: VanillaVAE.reparameterize draws eps itself; pin only its KLD/loss arithmetic
vv = models.VanillaVAE(3, 10)
torch.manual_seed(0)
mu2, lv2 = torch.randn(16, 10), torch.randn(16, 10)
rec, inp = torch.randn(16, 3, 8, 8), torch.randn(16, 3, 8, 8)
r = vv.loss_function(rec, inp, mu2, lv2, M_N=0.005)
save("vanilla_loss", mu=mu2, logvar=lv2, recons=rec, input=inp, M_N=0.005, loss=r["loss"],
     recons_loss=r["Reconstruction_Loss"], KLD=r["KLD"])
