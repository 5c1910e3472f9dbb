"""Mint the CT-codec golden vectors from the UNMODIFIED reference (models/ct_mcq_vae.py:306-311, 472-496), imported
live from /root/reference.  ``CTMCQVAE`` itself cannot be constructed here (its CausalTransition layer needs
torch_geometric's GATv2Conv, absent and not installable), but the three methods only read ``self.num_embeddings`` and
``self.codebooks``, so they are called unbound on a stand-in ``self``.

    python tests/golden/make_golden_ct.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_live  # noqa: E402

torch.set_num_threads(1)
models = ref_live.load()
from models.ct_mcq_vae import CTMCQVAE, CausalTransition  # noqa: E402


def save(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


def case(name, seed, B, C, H, W, N):
    torch.manual_seed(seed)
    me = types.SimpleNamespace(num_embeddings=N, codebooks=C)
    shape = [B, C * 32, H, W]                       # latents_shape: only entries 0, 2, 3 are read
    inds = torch.randint(0, N, (B, C, H, W))
    onehot = CTMCQVAE.ct_preprocess(me, inds, shape)              # [B, N, C*H, W]
    scores = torch.rand(B, N, C * H, W)
    scores[0, :, 0, 0] = 0.25                                     # an exact tie over all classes: first index wins
    post = CTMCQVAE.ct_postprocess(me, scores, shape)             # [B, C, H, W]
    rt = CTMCQVAE.ct_postprocess(me, onehot, shape)               # round trip
    assert torch.equal(rt, inds)
    latent = (torch.rand(B, N, C * H, W) * 0.2).requires_grad_(True)
    with torch.no_grad():
        latent[:, : N // 2] *= 1e-4                               # plenty of values under the 1e-4 clamp
    latent_y = torch.rand(B, N, C * H, W)
    loss = CausalTransition.latent_CrossEntropy_loss(me, latent, latent_y)
    (1.7 * loss).backward()
    save(name, B=B, C=C, H=H, W=W, N=N, inds=inds, onehot=onehot.contiguous(), scores=scores, post=post, latent=latent,
         latent_y=latent_y, ce=loss, g_ce=1.7, g_latent=latent.grad)


case("ct_codec_cfg3", 1250, 3, 1, 8, 8, 64)       # configs/ct_mcq_vae.yaml geometry (1 codebook, K=64, 8x8), 3 images
case("ct_codec_mcq", 1320, 2, 4, 8, 8, 32)        # 4 codebooks
case("ct_codec_odd", 7, 3, 2, 5, 3, 10)           # S not a multiple of 4: scalar path
