"""Autograd bindings of the quantiser kernels (thin: shapes, pointers, stream; all maths is in libctvq.so)."""
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

Tensor = torch.Tensor


def _shape(latents: Tensor, codebooks: Sequence[Tensor], chan_stride: int):
    if latents.dim() != 4:
        raise RuntimeError(f"latents must be [B, D, H, W], got {tuple(latents.shape)}")
    b, dtot, h, w = latents.shape
    k, d = codebooks[0].shape
    c = len(codebooks)
    for e in codebooks:
        if tuple(e.shape) != (k, d):
            raise RuntimeError("all codebooks must share one [K, d] shape")
        if e.dtype != torch.float32 or not e.is_contiguous():
            raise RuntimeError("codebooks must be contiguous float32")
    if (c - 1) * chan_stride + d > dtot:
        # same failure class as the reference: a size mismatch inside torch.matmul (models/mcq_vae.py:33)
        raise RuntimeError(f"codebook slices ({c} x {d} channels, stride {chan_stride}) exceed the {dtot} latent channels")
    if latents.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"latents must be float32 (the reference's arithmetic type) or bfloat16, got {latents.dtype}")
    return b, dtot, h, w, c, d, k


def _dtype_code(t: Tensor) -> int:
    """bf16 mode (not defined by the reference -- `.bfloat16()` raises at models/vq_vae.py:43): the fp32 arithmetic contract
    applied to bf16-ROUNDED latents and codebooks, NATIVE in the kernels: latents, outputs and their
    gradients cross HBM as bf16 (half the bytes), the fp32 codebook parameters are rounded to bf16 as they are read,
    every product and sum is fp32, indices stay int64, losses and codebook gradients stay fp32."""
    return _lib.BF16 if t.dtype == torch.bfloat16 else _lib.F32


def _bf16_via_fp32(b, dtot, hw, c, d, k, cs) -> bool:
    """bf16 latents run NATIVELY (dtype = CTVQ_BF16: bf16 TMA slabs, tcgen05.mma.kind::f16, bf16 streams in the backward)
    on the multi-codebook shape of configs/mcq_vae.yaml, and natively on the generic kernels for small problems.  Large
    problems of any OTHER shape have no 16-bit tensor-core kernel yet: they are up-cast once and take the fp32 tcgen05
    kernels (bf16 values are exact in fp32 and the codebook is rounded first, so the result is the same contract), which is
    10x faster there than the generic bf16 kernels despite the two cast passes."""
    native_shape = d == 32 and cs == 1 and k <= 64 and ((c == 4 and dtot == 128 and hw in (64, 256)) or (c == 2 and dtot == 64 and hw == 64))
    return (not native_shape) and b * hw * c * k * d > (1 << 24)


def _counter_ptr(counter: Optional[Tensor], dev) -> Optional[int]:
    if counter is None:
        return None
    if counter.dtype != torch.int64 or counter.numel() != 1 or counter.device != dev:
        raise RuntimeError("near-tie counter must be ONE int64 element on the latents' device")
    return counter.data_ptr()


def new_near_tie_counter(device) -> Tensor:
    """Zero-initialised device counter the kernels ADD near-tie rows to (include/ctvq.h, CTVQ_NEAR_TIE_REL)."""
    return torch.zeros(1, dtype=torch.int64, device=device)


def compute_inds(latents_list: Sequence[Tensor], codebooks: Sequence[Tensor], chan_stride: int = 1,
                 counter: Optional[Tensor] = None) -> List[Tensor]:
    """argmin indices [B, C, H, W] int64 for each input tensor, ONE launch for up to 4 same-shape inputs.
    ``counter`` (one int64 on the device) accumulates the number of near-tie rows (north_star: relative top-2 gap
    below 1e-6, counted and reported).

    Replaces models/mcq_vae.py:26-39 / :100-110 (and the x / y pair of models/ct_mcq_vae.py:530,536)."""
    _lib.require_cuda(*latents_list, *codebooks)
    zs = [z.detach().contiguous() for z in latents_list]
    es = [e.detach() for e in codebooks]
    b, dtot, h, w, c, d, k = _shape(zs[0], es, chan_stride)
    for z in zs[1:]:
        if z.shape != zs[0].shape or z.dtype != zs[0].dtype:
            raise RuntimeError("paired inputs must share a shape and dtype")
    dev = zs[0].device
    outs = [torch.empty((b, c, h, w), dtype=torch.int64, device=dev) for _ in zs]
    if b == 0 or h * w == 0:
        return outs
    if zs[0].dtype == torch.bfloat16 and _bf16_via_fp32(b, dtot, h * w, c, d, k, chan_stride):
        zs = [z.float() for z in zs]
        es = [e.to(torch.bfloat16).float() for e in es]
    sp = _lib.stream_ptr(dev)
    ws = _lib.workspace(dev, sp, c, k, d)
    rc = _lib.lib().ctvq_argmin(_lib.ptr_array(zs), len(zs), _lib.ptr_array(es), b, dtot, h * w, c, d, k, chan_stride,
                                _dtype_code(zs[0]), _lib.ptr_array(outs), _counter_ptr(counter, dev), ws.data_ptr(), ws.numel(),
                                dev.index, sp)
    _lib.check(rc, "ctvq_argmin")
    return outs


class _Quantize(torch.autograd.Function):
    """(latents, codebooks[, indices]) -> (straight-through output, summed vq_loss, indices, per-codebook losses)."""

    @staticmethod
    def forward(ctx, latents: Tensor, beta: float, chan_stride: int, given_inds: Optional[Tensor], comm, counter,
                *codebooks):
        _lib.require_cuda(latents, *codebooks)
        z = latents.detach().contiguous()
        es = [e.detach() for e in codebooks]
        b, dtot, h, w, c, d, k = _shape(z, es, chan_stride)
        io_dtype = z.dtype
        dev = z.device
        if b == 0 or h * w == 0:
            # empty batch: the reference returns an empty tensor and mse_loss(empty) = nan (models/vq_vae.py:47-50); its
            # codebook gradient is one_hot[0,K]^T @ g[0,d] = zeros, which this rank must still CONTRIBUTE to the
            # all-reduce (a rank that skipped the collective would hang its peers)
            empty_inds = torch.empty((b, c, h, w), dtype=torch.int64, device=dev)
            nan = torch.full((), float("nan"), device=dev)
            ctx.empty = True
            ctx.meta = (float(beta), int(chan_stride), b, dtot, h, w, c, d, k, comm)
            ctx.dev, ctx.io_dtype = dev, io_dtype
            per = torch.full((c,), float("nan"), device=dev)
            ctx.mark_non_differentiable(empty_inds, per)
            ctx.set_materialize_grads(False)
            return torch.empty((b, c * d, h, w), dtype=io_dtype, device=dev), nan, empty_inds, per
        ctx.empty = False
        ctx.cast = io_dtype == torch.bfloat16 and _bf16_via_fp32(b, dtot, h * w, c, d, k, chan_stride)
        if ctx.cast:
            z = z.float()
            es = [e.to(torch.bfloat16).float() for e in es]
        dt = _dtype_code(z)
        out = torch.empty((b, c * d, h, w), dtype=z.dtype, device=dev)
        losses = torch.empty(c + 1, dtype=torch.float32, device=dev)
        sp = _lib.stream_ptr(dev)
        ws = _lib.workspace(dev, sp, c, k, d)
        L = _lib.lib()
        if given_inds is None:
            inds = torch.empty((b, c, h, w), dtype=torch.int64, device=dev)
            rc = L.ctvq_forward(z.data_ptr(), _lib.ptr_array(es), b, dtot, h * w, c, d, k, chan_stride, dt,
                                float(beta), inds.data_ptr(), out.data_ptr(), losses.data_ptr(),
                                _counter_ptr(counter, dev), ws.data_ptr(), ws.numel(), dev.index, sp)
            _lib.check(rc, "ctvq_forward")
        else:
            _lib.require_cuda(given_inds)
            if given_inds.numel() != b * c * h * w:
                raise RuntimeError(f"indices have {given_inds.numel()} elements, expected {b * c * h * w}")
            inds = given_inds.detach().to(torch.int64).reshape(b, c, h, w).contiguous()
            rc = L.ctvq_gather_st_loss(z.data_ptr(), _lib.ptr_array(es), inds.data_ptr(), b, dtot, h * w, c, d, k,
                                       chan_stride, dt, float(beta), out.data_ptr(), losses.data_ptr(),
                                       ws.data_ptr(), ws.numel(), dev.index, sp)
            _lib.check(rc, "ctvq_gather_st_loss")
            _lib.maybe_validate(ws, dev, sp, "compute_latents")
        ctx.save_for_backward(z, inds, *es)
        ctx.meta = (float(beta), int(chan_stride), b, dtot, h, w, c, d, k, comm)
        ctx.io_dtype = io_dtype
        if ctx.cast:
            out = out.to(torch.bfloat16)
        per = losses[:c]
        ctx.mark_non_differentiable(inds, per)
        ctx.set_materialize_grads(False)  # unused outputs arrive as None: no zero-fill kernels, no host sync
        return out, losses[c], inds, per

    @staticmethod
    def backward(ctx, g_out, g_loss, _g_inds, _g_per):
        if ctx.empty:
            beta, cs, b, dtot, h, w, c, d, k, comm = ctx.meta
            dev = ctx.dev
            ge = torch.zeros((c, k, d), dtype=torch.float32, device=dev)
            if comm is not None:
                ge = comm.allreduce_(ge)
            gz = torch.zeros((b, dtot, h, w), dtype=ctx.io_dtype, device=dev)
            return (gz, None, None, None, None, None, *ge.unbind(0))
        z, inds, *es = ctx.saved_tensors
        beta, cs, b, dtot, h, w, c, d, k, comm = ctx.meta
        dev = z.device
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=dev)
        g_loss = g_loss.to(torch.float32).contiguous()
        go_ptr = None
        dt = _dtype_code(z)
        if g_out is not None:
            g_out = g_out.contiguous().to(z.dtype)
            go_ptr = g_out.data_ptr()
        gz = torch.empty_like(z)
        ge = torch.empty((c, k, d), dtype=torch.float32, device=dev)
        sp = _lib.stream_ptr(dev)
        ws = _lib.workspace(dev, sp, c, k, d)
        if comm is not None and getattr(comm, "fuses_backward", False):
            # backward + the path's one collective in ONE launch: the kernel's last CTA all-reduces grad_E over NVLink
            table, world, rank, cmax, epoch, scale = comm.peer_args(c * k * d)
            red = torch.empty((c, k, d), dtype=torch.float32, device=dev)
            rc = _lib.lib().ctvq_backward_allreduce(z.data_ptr(), _lib.ptr_array(es), inds.data_ptr(), go_ptr,
                                                    g_loss.data_ptr(), b, dtot, h * w, c, d, k, cs, dt, beta,
                                                    gz.data_ptr(), ge.data_ptr(), table, world, rank, cmax, epoch, scale,
                                                    red.data_ptr(), ws.data_ptr(), ws.numel(), dev.index, sp)
            _lib.check(rc, "ctvq_backward_allreduce")
            ge = red
            comm = None
        else:
            rc = _lib.lib().ctvq_backward(z.data_ptr(), _lib.ptr_array(es), inds.data_ptr(), go_ptr, g_loss.data_ptr(),
                                          b, dtot, h * w, c, d, k, cs, dt, beta, gz.data_ptr(), ge.data_ptr(),
                                          ws.data_ptr(), ws.numel(), dev.index, sp)
            _lib.check(rc, "ctvq_backward")
        _lib.maybe_validate(ws, dev, sp, "quantiser backward")
        if comm is not None:
            ge = comm.allreduce_(ge)  # the one collective of the path (NCCL), on the backward kernel's stream
        if gz.dtype != ctx.io_dtype:
            gz = gz.to(ctx.io_dtype)
        return (gz, None, None, None, None, None, *ge.unbind(0))


def quantize(latents: Tensor, codebooks: Sequence[Tensor], beta: float, chan_stride: int = 1,
             inds: Optional[Tensor] = None, comm=None, counter: Optional[Tensor] = None
             ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Fused forward (argmin + gather + loss + straight-through), or gather by the given ``inds``.
    ``counter``: optional one-element int64 device tensor the kernel adds this call's near-tie rows to.

    Replaces models/vq_vae.py:24-55 and models/mcq_vae.py:41-64,112-137."""
    return _Quantize.apply(latents, beta, chan_stride, inds, comm, counter, *codebooks)
