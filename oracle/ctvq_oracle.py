"""CPU oracle for the codebook-quantiser hot path — TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The package ``ct_vae_b200`` never does: its ops raise when the CUDA
library is missing instead of falling back to anything in here.

It is an independent restatement (not a copy) of the reference's arithmetic, written against the same
ATen CPU operators so that on identical inputs it reproduces the reference bit for bit.  Parity status:
**pinned** — ``tests/golden/make_golden.py`` runs the UNMODIFIED reference (imported live from
``/root/reference`` through ``oracle/ref_live.py``) and stores its outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files, and
``tests/test_oracle_vs_reference.py`` re-checks against the live import whenever the tree is mounted.
(The reference's own tests hold no golden vectors for this path: ``tests/test_vq_vae.py:17-29`` only
prints.)

All tensors are torch CPU tensors.  ``latents`` is NCHW ``[B, D, H, W]``; codebooks are ``[K, d]``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

Tensor = torch.Tensor

NEAR_TIE_REL = 1e-6  # north_star: "relative top-2 distance gap below 1e-6" is a near-tie


# ----------------------------------------------------------------------------------------------
# single codebook  (reference: models/vq_vae.py:24-55, models/mcq_vae.py:26-64)
# ----------------------------------------------------------------------------------------------
def _rows(latents: Tensor) -> Tensor:
    """NCHW -> [B*H*W, D] rows (reference: vq_vae.py:25-27 / mcq_vae.py:27-29)."""
    b, d, h, w = latents.shape
    return latents.permute(0, 2, 3, 1).contiguous().view(b * h * w, d)


def distances(rows: Tensor, codebook: Tensor) -> Tensor:
    """``(|z|^2 + |e|^2) - 2 z.e^T`` in exactly the reference's association (vq_vae.py:30-32).

    Python precedence makes the reference's expression ``(a + b) - c``; the order matters for which
    code wins a near-tie, so it is kept.
    """
    zz = torch.sum(rows ** 2, dim=1, keepdim=True)
    ee = torch.sum(codebook ** 2, dim=1)
    cross = torch.matmul(rows, codebook.t())
    return zz + ee - 2 * cross


def vq_compute_inds(latents: Tensor, codebook: Tensor) -> Tensor:
    """argmin over codes, first minimum wins (vq_vae.py:35, mcq_vae.py:37-39). -> [B,H,W] int64."""
    b, _, h, w = latents.shape
    return torch.argmin(distances(_rows(latents), codebook), dim=1).view(b, h, w)


def vq_compute_latents(latents: Tensor, inds: Tensor, codebook: Tensor, beta: float) -> Tuple[Tensor, Tensor]:
    """Gather + losses + straight-through output (vq_vae.py:38-55, mcq_vae.py:41-64).

    The reference gathers with ``one_hot @ E``; a product of a one-hot row with E is the selected row
    exactly, so an index-select is bit-identical.  ``F.mse_loss`` is ``mean((a-b)^2)``; both loss terms
    have the same value, combined as ``m*beta + m`` (vq_vae.py:50).
    """
    b, d, h, w = latents.shape
    z = latents.permute(0, 2, 3, 1).contiguous()
    q = codebook.index_select(0, inds.reshape(-1)).view(b, h, w, d)
    m = torch.mean((q - z) ** 2)
    loss = m * beta + m
    out = z + (q - z)  # two roundings; NOT bit-equal to q (vq_vae.py:53)
    return out.permute(0, 3, 1, 2).contiguous(), loss


def vq_forward(latents: Tensor, codebook: Tensor, beta: float) -> Tuple[Tensor, Tensor, Tensor]:
    inds = vq_compute_inds(latents, codebook)
    out, loss = vq_compute_latents(latents, inds, codebook, beta)
    return out, loss, inds


def vq_backward(latents: Tensor, inds: Tensor, codebook: Tensor, beta: float,
                g_out: Tensor, g_loss: Tensor) -> Tuple[Tensor, Tensor]:
    """Explicit gradients of ``vq_compute_latents`` (autograd of vq_vae.py:43-53).

    grad_z = g_out + g_loss * beta * 2 (z - q) / (N d)         (commitment term; ST passes g_out)
    grad_E[k] = g_loss * sum_{n: idx_n = k} 2 (q_n - z_n) / (N d)   (embedding term only)
    """
    b, d, h, w = latents.shape
    n_el = b * h * w * d
    z = _rows(latents)
    q = codebook.index_select(0, inds.reshape(-1))
    diff = q - z
    gz_rows = (-2.0 * beta / n_el) * g_loss * diff
    gz = g_out + gz_rows.view(b, h, w, d).permute(0, 3, 1, 2)
    ge = torch.zeros_like(codebook)
    ge.index_add_(0, inds.reshape(-1), (2.0 / n_el) * g_loss * diff)
    return gz.contiguous(), ge


# ----------------------------------------------------------------------------------------------
# multiple codebooks  (reference: models/mcq_vae.py:100-137)
# ----------------------------------------------------------------------------------------------
def mcq_slice(latents: Tensor, i: int, d: int, chan_stride: int = 1) -> Tensor:
    """Codebook i reads channels ``i*chan_stride ... +d`` — the reference slices ``[:, i:i+d]``
    (mcq_vae.py:104,117), i.e. chan_stride = 1 and the slices OVERLAP."""
    return latents[:, i * chan_stride:i * chan_stride + d, :, :]


def mcq_compute_inds(latents: Tensor, codebooks: Sequence[Tensor], chan_stride: int = 1) -> Tensor:
    d = codebooks[0].shape[1]
    per = [vq_compute_inds(mcq_slice(latents, i, d, chan_stride), e) for i, e in enumerate(codebooks)]
    return torch.stack(per, 1)  # [B, C, H, W]  (mcq_vae.py:108)


def mcq_compute_latents(latents: Tensor, inds: Tensor, codebooks: Sequence[Tensor], beta: float,
                        chan_stride: int = 1) -> Tuple[Tensor, Tensor, List[Tensor]]:
    d = codebooks[0].shape[1]
    outs, losses = [], []
    for i, e in enumerate(codebooks):
        o, l = vq_compute_latents(mcq_slice(latents, i, d, chan_stride), inds[:, i], e, beta)
        outs.append(o)
        losses.append(l)
    total = sum(losses)  # python sum: ((0 + l0) + l1) + ...   (mcq_vae.py:125)
    return torch.cat(outs, 1), total, losses


def mcq_forward(latents: Tensor, codebooks: Sequence[Tensor], beta: float, chan_stride: int = 1):
    inds = mcq_compute_inds(latents, codebooks, chan_stride)
    out, total, losses = mcq_compute_latents(latents, inds, codebooks, beta, chan_stride)
    return out, total, inds, losses


def mcq_backward(latents: Tensor, inds: Tensor, codebooks: Sequence[Tensor], beta: float,
                 g_out: Tensor, g_loss: Tensor, chan_stride: int = 1) -> Tuple[Tensor, List[Tensor]]:
    """grad wrt the FULL latents tensor: overlapping slices accumulate (SURVEY §7 hard part 4);
    channels no slice touches get zero."""
    d = codebooks[0].shape[1]
    gz = torch.zeros_like(latents)
    ges = []
    for i, e in enumerate(codebooks):
        gsub, ge = vq_backward(mcq_slice(latents, i, d, chan_stride), inds[:, i], e, beta,
                               g_out[:, i * d:(i + 1) * d], g_loss)
        gz[:, i * chan_stride:i * chan_stride + d] += gsub
        ges.append(ge)
    return gz, ges


# ----------------------------------------------------------------------------------------------
# near-tie accounting
# ----------------------------------------------------------------------------------------------
def classify_index_mismatches(latents: Tensor, codebook: Tensor, idx_a: Tensor, idx_b: Tensor,
                              rel: float = NEAR_TIE_REL) -> Tuple[int, int]:
    """Compare two index maps.  Returns (near_tie_mismatches, hard_mismatches).

    A mismatch is a *near-tie* when the two chosen codes' float64 distances differ by less than
    ``rel`` times the operand scale ``|z|^2 + |e|^2`` of the reference formula (the quantity whose
    fp32 rounding sets the resolution of the comparison; for untrained data it equals the distance
    itself to within a few percent).  Anything else is a hard mismatch = a parity failure.
    """
    rows = _rows(latents).double()
    e = codebook.double()
    a = idx_a.reshape(-1)
    b = idx_b.reshape(-1)
    bad = (a != b).nonzero().flatten()
    if bad.numel() == 0:
        return 0, 0
    zr = rows[bad]
    ea, eb = e[a[bad]], e[b[bad]]
    da = ((zr - ea) ** 2).sum(1)
    db = ((zr - eb) ** 2).sum(1)
    scale = (zr ** 2).sum(1) + torch.maximum((ea ** 2).sum(1), (eb ** 2).sum(1))
    near = (da - db).abs() < rel * scale
    return int(near.sum()), int((~near).sum())


def count_near_tie_rows(latents: Tensor, codebook: Tensor, rel: float = NEAR_TIE_REL) -> int:
    """Rows whose best and second-best fp32 distances are within ``rel`` (relative)."""
    dist = distances(_rows(latents), codebook)
    if dist.shape[1] < 2:
        return 0
    top2 = torch.topk(dist, 2, dim=1, largest=False).values
    gap = (top2[:, 1] - top2[:, 0]).abs()
    return int((gap <= rel * top2[:, 0].abs()).sum())  # '<=': an exact tie at distance 0 counts too (include/ctvq.h)


# ----------------------------------------------------------------------------------------------
# Gaussian branch  (reference: models/vanilla_vae.py:107-117,143; models/beta_vae.py:112-122,141)
# ----------------------------------------------------------------------------------------------
def reparameterize(mu: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """``eps * exp(0.5 logvar) + mu`` with eps SUPPLIED (the reference draws it with randn_like,
    vanilla_vae.py:116; parity is defined on a given eps)."""
    std = torch.exp(0.5 * logvar)
    return eps * std + mu


def kld(mu: Tensor, logvar: Tensor) -> Tensor:
    """``mean_b(-0.5 * sum_d(1 + lv - mu^2 - e^lv))`` (vanilla_vae.py:143, beta_vae.py:141)."""
    return torch.mean(-0.5 * torch.sum(1 + logvar - mu ** 2 - logvar.exp(), dim=1), dim=0)


def reparam_kld_backward(mu: Tensor, logvar: Tensor, eps: Tensor, g_z: Tensor, g_kld: Tensor):
    """Explicit gradients of (reparameterize, kld) wrt (mu, logvar)."""
    b = mu.shape[0]
    g_mu = g_z + g_kld * mu / b
    g_lv = g_z * eps * 0.5 * torch.exp(0.5 * logvar) + g_kld * (-0.5 / b) * (1 - logvar.exp())
    return g_mu, g_lv


# ---- CT-mode codec: restated on the same ATen CPU operators as the reference ------------------
def ct_preprocess(x: Tensor, latents_shape: Sequence[int], num_embeddings: int, codebooks: int) -> Tensor:
    """models/ct_mcq_vae.py:472-483: one-hot of the code indices, [B,K,H,W] -> [B,N,K*H,W]."""
    oh = torch.nn.functional.one_hot(x, num_classes=num_embeddings).to(dtype=torch.float32)
    oh = oh.view((latents_shape[0], codebooks * latents_shape[2], latents_shape[3], num_embeddings))
    return oh.permute(0, 3, 1, 2)


def ct_postprocess(x: Tensor, latents_shape: Sequence[int], num_embeddings: int, codebooks: int) -> Tensor:
    """models/ct_mcq_vae.py:485-496: argmax over the class dimension, [B,N,K*H,W] -> [B,K,H,W]."""
    y = x.permute(0, 2, 3, 1)
    y = y.reshape((latents_shape[0], codebooks, latents_shape[2], latents_shape[3], num_embeddings))
    return torch.argmax(y, dim=-1)


def latent_cross_entropy_loss(latent: Tensor, latent_y: Tensor) -> Tensor:
    """models/ct_mcq_vae.py:306-311: cross-entropy of log(clamp(latent, 1e-4)) against argmax(latent_y)."""
    a = latent.permute(0, 2, 3, 1).reshape((-1, latent.size(1)))
    a = a.clamp(min=1e-4).log()
    t = latent_y.permute(0, 2, 3, 1).reshape((-1, latent_y.size(1)))
    t = torch.argmax(t, dim=-1)
    return torch.nn.functional.cross_entropy(a, t)
