"""ctypes wrapper of the plain-C oracle (oracle/ctvq_oracle_c.c) — TEST INFRASTRUCTURE ONLY.

The C oracle evaluates every sum in the order the CUDA kernels are specified to use, so GPU results
can be compared bit-for-bit on every row (no near-tie allowance).  Build: ``make -C oracle``.
"""
import ctypes
import os
import subprocess
from typing import Sequence, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    subprocess.run(["make", "-C", _HERE, "libctvq_oracle.so"], check=True, capture_output=True)
    return os.path.join(_HERE, "libctvq_oracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libctvq_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _f(t):
    return ctypes.c_void_p(t.data_ptr())


def _ptrs(codebooks: Sequence[torch.Tensor]):
    return (ctypes.c_void_p * len(codebooks))(*[c.data_ptr() for c in codebooks])


def _prep(latents, codebooks):
    z = latents.detach().contiguous().float()
    es = [e.detach().contiguous().float() for e in codebooks]
    b, dtot, h, w = z.shape
    k, d = es[0].shape
    return z, es, b, dtot, h * w, len(es), d, k


def argmin(latents: torch.Tensor, codebooks: Sequence[torch.Tensor], chan_stride: int = 1) -> torch.Tensor:
    z, es, b, dtot, hw, c, d, k = _prep(latents, codebooks)
    out = torch.empty(b, c, latents.shape[2], latents.shape[3], dtype=torch.int64)
    lib().ctvq_c_argmin(_f(z), _ptrs(es), ctypes.c_int64(b), dtot, hw, c, d, k, chan_stride, _f(out))
    return out


def neartie_count(latents: torch.Tensor, codebooks: Sequence[torch.Tensor], chan_stride: int = 1) -> int:
    """(row, codebook) pairs with a relative top-2 distance gap below 1e-6 (north_star), kernels' evaluation order."""
    z, es, b, dtot, hw, c, d, k = _prep(latents, codebooks)
    fn = lib().ctvq_c_neartie_count
    fn.restype = ctypes.c_int64
    return int(fn(_f(z), _ptrs(es), ctypes.c_int64(b), dtot, hw, c, d, k, chan_stride))


def gather_st_loss(latents, inds, codebooks, beta: float, chan_stride: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    z, es, b, dtot, hw, c, d, k = _prep(latents, codebooks)
    idx = inds.reshape(b, c, hw).contiguous()
    q = torch.empty(b, c * d, latents.shape[2], latents.shape[3])
    loss = torch.empty(c + 1)
    lib().ctvq_c_gather_st_loss(_f(z), _ptrs(es), _f(idx), ctypes.c_int64(b), dtot, hw, c, d, k, chan_stride,
                                ctypes.c_float(beta), _f(q), _f(loss))
    return q, loss


def backward(latents, inds, codebooks, beta: float, g_out, g_loss: float, chan_stride: int = 1):
    z, es, b, dtot, hw, c, d, k = _prep(latents, codebooks)
    idx = inds.reshape(b, c, hw).contiguous()
    go = g_out.contiguous().float()
    gz = torch.empty_like(z)
    ge = torch.empty(c, k, d)
    lib().ctvq_c_backward(_f(z), _ptrs(es), _f(idx), _f(go), ctypes.c_float(g_loss), ctypes.c_int64(b), dtot, hw, c,
                          d, k, chan_stride, ctypes.c_float(beta), _f(gz), _f(ge))
    return gz, ge


def reparam_kld(mu, logvar, eps):
    mu, logvar, eps = (t.contiguous().float() for t in (mu, logvar, eps))
    z = torch.empty_like(mu)
    k = torch.empty(1)
    lib().ctvq_c_reparam_kld(_f(mu), _f(logvar), _f(eps), ctypes.c_int64(mu.shape[0]), mu.shape[1], _f(z), _f(k))
    return z, k[0]


# ---- CT-mode codec (models/ct_mcq_vae.py:472-496, 306-311): tensors as [B, K, S] / [B, S] like the kernels -----------
def ct_onehot(inds: torch.Tensor, num_embeddings: int) -> torch.Tensor:
    """[B, C, H, W] int64 -> one-hot fp32 [B, N, C*H, W] (contiguous)."""
    b, c, h, w = inds.shape
    idx = inds.contiguous()
    out = torch.empty(b, num_embeddings, c * h, w)
    lib().ctvq_c_onehot(_f(idx), ctypes.c_int64(b), ctypes.c_int64(c * h * w), num_embeddings, _f(out))
    return out


def ct_class_argmax(scores: torch.Tensor, codebooks: int) -> torch.Tensor:
    """[B, N, C*H, W] fp32 -> [B, C, H, W] int64."""
    x = scores.detach().contiguous().float()
    b, n, ch, w = x.shape
    out = torch.empty(b, codebooks, ch // codebooks, w, dtype=torch.int64)
    lib().ctvq_c_class_argmax(_f(x), ctypes.c_int64(b), ctypes.c_int64(ch * w), n, _f(out))
    return out


def ct_latent_ce(latent: torch.Tensor, latent_y: torch.Tensor, g: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loss, d loss*g / d latent) of latent_CrossEntropy_loss."""
    x, y = latent.detach().contiguous().float(), latent_y.detach().contiguous().float()
    b, n, ch, w = x.shape
    loss = torch.empty(())
    gx = torch.empty_like(x)
    lib().ctvq_c_latent_ce(_f(x), _f(y), ctypes.c_int64(b), ctypes.c_int64(ch * w), n, ctypes.c_float(g), _f(loss), _f(gx))
    return loss, gx
