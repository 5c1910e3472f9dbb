"""CT-mode codec: the index <-> one-hot converters and the one-hot cross-entropy that sit either side of the quantiser
in ``CTMCQVAE``.

    reference method (models/ct_mcq_vae.py)                     here
    ------------------------------------------------------     -----------------------------------------
    CTMCQVAE.ct_preprocess(x, latents_shape)      :472-483      ct_preprocess(x, latents_shape, num_embeddings, codebooks)
    CTMCQVAE.ct_postprocess(x, latents_shape)     :485-496      ct_postprocess(x, latents_shape, num_embeddings, codebooks)
    CausalTransition.latent_CrossEntropy_loss(l, l_y)     :306-311      latent_cross_entropy_loss(latent, latent_y)

Same argument meaning, shapes, dtypes and values.  ``ct_preprocess`` returns the one-hot as a CONTIGUOUS
``[B, N, K*H, W]`` tensor (the reference returns a permuted view of ``[B, K*H, W, N]`` with the same values).
``install(cls)`` rebinds the three methods on a ``CTMCQVAE`` class.  CUDA tensors only: there is no CPU fallback.
"""
from typing import Sequence

import torch

from . import _lib

Tensor = torch.Tensor


def ct_preprocess(x: Tensor, latents_shape: Sequence[int], num_embeddings: int, codebooks: int) -> Tensor:
    """[B, K, H, W] int64 code indices -> one-hot fp32 [B, N, K*H, W] (models/ct_mcq_vae.py:472-483)."""
    _lib.require_cuda(x)
    if x.dtype != torch.int64:
        raise RuntimeError("one_hot is only applicable to index tensor of type LongTensor.")  # F.one_hot's own message
    b, h, w = int(latents_shape[0]), int(latents_shape[2]), int(latents_shape[3])
    idx = x.contiguous()
    s = codebooks * h * w
    if idx.numel() != b * s:
        raise RuntimeError(f"shape '{[b, codebooks * h, w, num_embeddings]}' is invalid for input of size {idx.numel() * num_embeddings}")
    out = torch.empty((b, num_embeddings, codebooks * h, w), dtype=torch.float32, device=x.device)
    if b * s == 0:
        return out
    dev = x.device
    sp = _lib.stream_ptr(dev)
    ws = _lib.workspace(dev, sp)
    rc = _lib.lib().ctvq_onehot_from_inds(idx.data_ptr(), b, s, num_embeddings, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                          dev.index, sp)
    _lib.check(rc, "ctvq_onehot_from_inds")
    _lib.maybe_validate(ws, dev, sp, "ct_preprocess")  # CTVQ_VALIDATE / set_validate(True): raise like F.one_hot does
    return out


def ct_postprocess(x: Tensor, latents_shape: Sequence[int], num_embeddings: int, codebooks: int) -> Tensor:
    """[B, N, K*H, W] fp32 class scores -> argmax code indices [B, K, H, W] int64 (models/ct_mcq_vae.py:485-496)."""
    _lib.require_cuda(x)
    if x.dtype != torch.float32:
        raise RuntimeError("ct_postprocess expects float32 scores (the reference's one-hot dtype)")
    b, h, w = int(latents_shape[0]), int(latents_shape[2]), int(latents_shape[3])
    s = codebooks * h * w
    if x.dim() != 4 or x.shape[0] != b or x.shape[1] != num_embeddings or x.shape[2] * x.shape[3] != s:
        raise RuntimeError(f"shape '{[b, codebooks, h, w, num_embeddings]}' is invalid for input of size {x.numel()}")
    xc = x.detach().contiguous()
    out = torch.empty((b, codebooks, h, w), dtype=torch.int64, device=x.device)
    if b * s == 0:
        return out
    dev = x.device
    rc = _lib.lib().ctvq_inds_from_onehot(xc.data_ptr(), b, s, num_embeddings, out.data_ptr(), dev.index, _lib.stream_ptr(dev))
    _lib.check(rc, "ctvq_inds_from_onehot")
    return out


class _LatentCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, latent: Tensor, latent_y: Tensor):
        _lib.require_cuda(latent, latent_y)
        if latent.dim() != 4 or latent.shape != latent_y.shape:
            raise RuntimeError(f"latent / latent_y must share a [B, D, H, W] shape, got {tuple(latent.shape)}, {tuple(latent_y.shape)}")
        if latent.dtype != torch.float32 or latent_y.dtype != torch.float32:
            raise RuntimeError("latent_cross_entropy_loss is float32 (the reference's arithmetic type)")
        x, y = latent.detach().contiguous(), latent_y.detach().contiguous()
        b, k = x.shape[0], x.shape[1]
        s = x.shape[2] * x.shape[3]
        dev = x.device
        tgt = torch.empty((b, s), dtype=torch.int64, device=dev)
        rowsum = torch.empty((b, s), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        sp = _lib.stream_ptr(dev)
        ws = _lib.workspace(dev, sp)
        rc = _lib.lib().ctvq_latent_ce_fwd(x.data_ptr(), y.data_ptr(), b, s, k, tgt.data_ptr(), rowsum.data_ptr(), loss.data_ptr(),
                                           ws.data_ptr(), ws.numel(), dev.index, sp)
        _lib.check(rc, "ctvq_latent_ce_fwd")
        ctx.save_for_backward(x, tgt, rowsum)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        x, tgt, rowsum = ctx.saved_tensors
        dev = x.device
        g = g_loss.to(torch.float32).contiguous()
        gx = torch.empty_like(x)
        b, k = x.shape[0], x.shape[1]
        s = x.shape[2] * x.shape[3]
        rc = _lib.lib().ctvq_latent_ce_bwd(x.data_ptr(), tgt.data_ptr(), rowsum.data_ptr(), g.data_ptr(), b, s, k, gx.data_ptr(),
                                           dev.index, _lib.stream_ptr(dev))
        _lib.check(rc, "ctvq_latent_ce_bwd")
        return gx, None


def latent_cross_entropy_loss(latent: Tensor, latent_y: Tensor) -> Tensor:
    """mean cross-entropy of log(clamp(latent, 1e-4)) against argmax(latent_y) over dim 1 (models/ct_mcq_vae.py:306-311);
    the gradient flows to ``latent`` only (``latent_y`` is detached by the caller, :300)."""
    if latent.numel() == 0:
        return torch.full((), float("nan"), device=latent.device)  # F.cross_entropy of no rows
    return _LatentCE.apply(latent, latent_y)


def install(ctmcqvae_cls=None, causal_transition_cls=None) -> None:
    """Rebind ct_preprocess / ct_postprocess on the reference's ``CTMCQVAE`` class (they only read
    ``self.num_embeddings`` and ``self.codebooks``, models/ct_mcq_vae.py:480-481,494) and latent_CrossEntropy_loss on
    its ``CausalTransition`` class (models/ct_mcq_vae.py:306-311, called from latent_loss :299-301)."""
    if ctmcqvae_cls is not None:
        ctmcqvae_cls.ct_preprocess = lambda self, x, latents_shape: ct_preprocess(x, latents_shape, self.num_embeddings, self.codebooks)
        ctmcqvae_cls.ct_postprocess = lambda self, x, latents_shape: ct_postprocess(x, latents_shape, self.num_embeddings, self.codebooks)
    if causal_transition_cls is not None:
        causal_transition_cls.latent_CrossEntropy_loss = lambda self, latent, latent_y: latent_cross_entropy_loss(latent, latent_y)
