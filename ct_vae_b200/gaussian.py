"""Fused Gaussian branch: reparameterisation + KL term in one pass over (mu, logvar, eps).

Replaces ``VanillaVAE.reparameterize`` / ``BetaVAE.reparameterize`` (models/vanilla_vae.py:107-117,
models/beta_vae.py:112-122) and the KLD line of their ``loss_function`` (vanilla_vae.py:143,
beta_vae.py:141).  ``eps`` is drawn with ``torch.randn_like(logvar)`` — the same generator call the
reference makes on ``std`` (same shape/dtype/device), so a seeded run consumes the RNG stream identically.
"""
from typing import Optional, Tuple

import torch

from . import _lib

Tensor = torch.Tensor


class _ReparamKLD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu: Tensor, logvar: Tensor, eps: Tensor):
        _lib.require_cuda(mu, logvar, eps)
        if mu.dim() != 2 or mu.shape != logvar.shape or mu.shape != eps.shape:
            raise RuntimeError(f"mu/logvar/eps must share a [B, L] shape, got {tuple(mu.shape)}, "
                               f"{tuple(logvar.shape)}, {tuple(eps.shape)}")
        if mu.dtype != torch.float32:
            raise RuntimeError("the Gaussian branch is float32 (the reference's arithmetic type)")
        mu_c, lv_c, eps_c = mu.detach().contiguous(), logvar.detach().contiguous(), eps.detach().contiguous()
        dev = mu.device
        z = torch.empty_like(mu_c)
        kld = torch.empty((), dtype=torch.float32, device=dev)
        sp = _lib.stream_ptr(dev)
        ws = _lib.workspace(dev, sp)
        rc = _lib.lib().ctvq_reparam_kld_fwd(mu_c.data_ptr(), lv_c.data_ptr(), eps_c.data_ptr(), mu.shape[0],
                                             mu.shape[1], z.data_ptr(), kld.data_ptr(), ws.data_ptr(), ws.numel(),
                                             dev.index, sp)
        _lib.check(rc, "ctvq_reparam_kld_fwd")
        ctx.save_for_backward(mu_c, lv_c, eps_c)
        ctx.set_materialize_grads(False)
        return z, kld

    @staticmethod
    def backward(ctx, g_z, g_kld):
        mu, lv, eps = ctx.saved_tensors
        dev = mu.device
        g_mu = torch.empty_like(mu)
        g_lv = torch.empty_like(lv)
        gz_ptr = g_z.contiguous().data_ptr() if g_z is not None else None
        if g_z is not None:
            g_z = g_z.contiguous()
            gz_ptr = g_z.data_ptr()
        gk_ptr = None
        if g_kld is not None:
            g_kld = g_kld.to(torch.float32).contiguous()
            gk_ptr = g_kld.data_ptr()
        rc = _lib.lib().ctvq_reparam_kld_bwd(mu.data_ptr(), lv.data_ptr(), eps.data_ptr(), gz_ptr, gk_ptr, mu.shape[0],
                                             mu.shape[1], g_mu.data_ptr(), g_lv.data_ptr(), dev.index,
                                             _lib.stream_ptr(dev))
        _lib.check(rc, "ctvq_reparam_kld_bwd")
        return g_mu, g_lv, None


def reparam_kld(mu: Tensor, logvar: Tensor, eps: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """-> (z = eps*exp(0.5*logvar) + mu,  kld = mean_b(-0.5*sum(1 + logvar - mu^2 - exp(logvar))))."""
    if eps is None:
        eps = torch.randn_like(logvar)  # vanilla_vae.py:116 draws randn_like(std): same stream position
    return _ReparamKLD.apply(mu, logvar, eps)


def reparameterize(mu: Tensor, logvar: Tensor, eps: Optional[Tensor] = None) -> Tensor:
    return reparam_kld(mu, logvar, eps)[0]


class FusedGaussianMixin:
    """Mix into a VanillaVAE / BetaVAE-shaped model: ``reparameterize`` also produces the KL term, which
    ``loss_function`` then consumes instead of recomputing it (same dict keys and weighting as
    models/vanilla_vae.py:133-146 and models/beta_vae.py:131-152)."""

    _fused_kld = None

    def reparameterize(self, mu: Tensor, logvar: Tensor) -> Tensor:
        z, kld = reparam_kld(mu, logvar)
        self._fused_kld = (mu, logvar, kld)
        return z

    def _kld(self, mu: Tensor, log_var: Tensor) -> Tensor:
        cached = self._fused_kld
        if cached is not None and cached[0] is mu and cached[1] is log_var:
            return cached[2]
        return reparam_kld(mu, log_var, torch.zeros_like(mu))[1]


def vanilla_loss(recons: Tensor, inp: Tensor, kld: Tensor, M_N: float) -> dict:
    """models/vanilla_vae.py:139-146 given the fused KL term."""
    recons_loss = torch.nn.functional.mse_loss(recons, inp)
    loss = recons_loss + M_N * kld
    return {"loss": loss, "Reconstruction_Loss": recons_loss.detach(), "KLD": -kld.detach()}


def beta_loss(recons: Tensor, inp: Tensor, kld: Tensor, M_N: float, *, loss_type: str, beta: float, gamma: float,
              C_max: float, C_stop_iter: float, num_iter: int) -> dict:
    """models/beta_vae.py:139-152 given the fused KL term (``num_iter`` already incremented, :132)."""
    recons_loss = torch.nn.functional.mse_loss(recons, inp)
    if loss_type == "H":
        loss = recons_loss + beta * M_N * kld
    elif loss_type == "B":
        C = min(max(C_max / C_stop_iter * num_iter, 0.0), C_max)
        loss = recons_loss + gamma * M_N * (kld - C).abs()
    else:
        raise ValueError("Undefined loss type.")
    return {"loss": loss, "Reconstruction_Loss": recons_loss, "KLD": kld}
