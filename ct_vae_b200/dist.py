"""The one collective of the path: all-reduce of the stacked codebook gradient [C, K, d] over NCCL
(NVLink 5 / NVSwitch), issued through the C ABI on the backward kernel's stream.

Replaces the share of Lightning's DDP bucket all-reduce that carries ``vq_layer.*.embedding.weight.grad``
(run.py:99).  ``torch.distributed`` is plumbing only: it broadcasts the 128-byte NCCL unique id.
On CPU (``gloo`` tests) the same class falls back to ``torch.distributed.all_reduce`` so host logic is
testable without a GPU; that path never runs a quantiser kernel.
"""
import ctypes
import glob
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib


def _find_libnccl() -> Optional[str]:
    try:
        import nvidia.nccl as pkg  # the torch-bundled wheel
        for base in list(getattr(pkg, "__path__", [])):
            hits = glob.glob(os.path.join(base, "lib", "libnccl.so*"))
            if hits:
                return sorted(hits)[0]
    except Exception:
        pass
    return None


class CodebookGradComm:
    """Sum codebook gradients over ranks and scale by 1/world (DDP's averaging of per-rank mean losses)."""

    def __init__(self, group=None, device: Optional[torch.device] = None, average: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (it carries the NCCL unique id)")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.scale = 1.0 / self.world if average else 1.0
        self.device = device
        self._comm = None
        self.launches = 0
        if device is not None and device.type == "cuda":
            L = _lib.lib()
            _lib.check(L.ctvq_nccl_load((_find_libnccl() or "libnccl.so.2").encode()), "ctvq_nccl_load")
            ident = ctypes.create_string_buffer(128)
            if self.rank == 0:
                _lib.check(L.ctvq_nccl_unique_id(ident), "ctvq_nccl_unique_id")
            box = [ident.raw if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            comm = ctypes.c_void_p()
            _lib.check(L.ctvq_nccl_comm_init(ctypes.byref(comm), self.world, self.rank, box[0], device.index),
                       "ctvq_nccl_comm_init")
            self._comm = comm

    def allreduce_(self, grad: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return grad
        if grad.is_cuda:
            if self._comm is None:
                raise RuntimeError("CodebookGradComm was built without a CUDA device")
            sp = _lib.stream_ptr(grad.device)
            rc = _lib.lib().ctvq_allreduce_codebook_grad(self._comm, grad.data_ptr(), grad.numel(), self.scale,
                                                         grad.device.index, sp)
            _lib.check(rc, "ctvq_allreduce_codebook_grad")
        else:  # host-logic tests (gloo)
            dist.all_reduce(grad, group=self.group)
            grad.mul_(self.scale)
        self.launches += 1
        return grad

    def close(self):
        if self._comm is not None:
            _lib.lib().ctvq_nccl_comm_destroy(self._comm)
            self._comm = None


def shard_batch(global_batch: int, world: int, rank: int):
    """Contiguous even split of the batch dimension (DistributedSampler-style, datasets/transition.py:173-176);
    the remainder goes to the lowest ranks.  -> (start, stop)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
