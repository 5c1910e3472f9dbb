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


# ----------------------------------------------------------------------------------------------------------------------
# injection into the reference's VanillaVAE / BetaVAE
# ----------------------------------------------------------------------------------------------------------------------
def _fused_reparameterize(self, mu: Tensor, logvar: Tensor) -> Tensor:
    """Replaces VanillaVAE.reparameterize (models/vanilla_vae.py:107-117) / BetaVAE.reparameterize
    (models/beta_vae.py:112-122): same eps draw (randn_like on a tensor of logvar's shape/dtype/device), one kernel that
    also leaves the KL term of loss_function behind, keyed by the identity of (mu, logvar)."""
    z, kld = reparam_kld(mu, logvar)
    self._ctvq_kld = (mu, logvar, kld)
    return z


def _kld_of(self, mu: Tensor, log_var: Tensor) -> Tensor:
    cached = getattr(self, "_ctvq_kld", None)
    if cached is not None and cached[0] is mu and cached[1] is log_var:
        self._ctvq_kld = None  # single use: the autograd graph of that forward is about to be consumed
        return cached[2]
    # loss_function called on tensors that did not come from this module's reparameterize: KL term alone
    return reparam_kld(mu, log_var, torch.zeros_like(mu))[1]


def _vanilla_loss_function(self, *args, **kwargs) -> dict:
    """Replaces VanillaVAE.loss_function (models/vanilla_vae.py:128-146): same arguments, same dict."""
    recons, inp, mu, log_var = args[0], args[1], args[2], args[3]
    kld_weight = kwargs["M_N"]
    recons_loss = torch.nn.functional.mse_loss(recons, inp)
    kld_loss = _kld_of(self, mu, log_var)
    loss = recons_loss + kld_weight * kld_loss
    return {"loss": loss, "Reconstruction_Loss": recons_loss.detach(), "KLD": -kld_loss.detach()}


def _beta_loss_function(self, *args, **kwargs) -> dict:
    """Replaces BetaVAE.loss_function (models/beta_vae.py:130-152): same arguments, same dict, same ``num_iter`` counter
    and capacity schedule C = clamp(C_max / C_stop_iter * num_iter, 0, C_max) for loss type 'B'."""
    self.num_iter += 1
    recons, inp, mu, log_var = args[0], args[1], args[2], args[3]
    kld_weight = kwargs["M_N"]
    recons_loss = torch.nn.functional.mse_loss(recons, inp)
    kld_loss = _kld_of(self, mu, log_var)
    if self.loss_type == "H":
        loss = recons_loss + self.beta * kld_weight * kld_loss
    elif self.loss_type == "B":
        self.C_max = self.C_max.to(inp.device)
        C = torch.clamp(self.C_max / self.C_stop_iter * self.num_iter, 0, self.C_max.data[0])
        loss = recons_loss + self.gamma * kld_weight * (kld_loss - C).abs()
    else:
        raise ValueError("Undefined loss type.")
    return {"loss": loss, "Reconstruction_Loss": recons_loss, "KLD": kld_loss}


def install(*classes) -> int:
    """Rebind ``reparameterize`` and ``loss_function`` of the given VanillaVAE / BetaVAE classes (or of the ``models``
    package when called as ``install(models)``) to the fused kernel path.  A class is treated as Beta-shaped when it has
    the ``num_iter`` counter of models/beta_vae.py:10.  Idempotent; returns the number of classes patched."""
    todo = []
    for c in classes:
        if isinstance(c, type):
            todo.append(c)
        else:  # a package / module: pick the two classes by name
            todo += [getattr(c, n) for n in ("VanillaVAE", "BetaVAE") if isinstance(getattr(c, n, None), type)]
    n = 0
    for cls in todo:
        if cls.__dict__.get("_ctvq_fused_gaussian", False):
            continue
        cls.reparameterize = _fused_reparameterize
        cls.loss_function = _beta_loss_function if hasattr(cls, "num_iter") else _vanilla_loss_function
        cls._ctvq_fused_gaussian = True
        n += 1
    return n
