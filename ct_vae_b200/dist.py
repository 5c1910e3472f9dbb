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


class PeerGradComm:
    """The same collective WITHOUT NCCL, fused into the backward kernel: a one-shot all-reduce over NVLink peer memory
    (ctvq_backward_allreduce / ctvq_peer_allreduce in include/ctvq.h, protocol in csrc/ctvq_peer.cuh).

    The LAST CTA of the backward kernel pushes the finished codebook gradient into every rank's symmetric receive buffer
    (posted NVLink stores), posts a flag, waits for the peers' flags on local memory and sums the `world` slots in rank
    order (bit-identical on every rank), scaled by 1/world.  One launch, on the caller's stream: the reduced gradient is
    ordinary stream-ordered data when backward returns -- no side stream, nothing for the caller to wait on, safe for
    AccumulateGrad / hooks / clip_grad_norm_ / optimizer.step.  torch.distributed only carries the 64-byte IPC handles.
    Single node, world <= 8 (every GPU reaches every peer through NVSwitch).

    A peer that does not arrive within CTVQ_PEER_TIMEOUT_MS (default 30 s) does not trap the GPU: the kernel finishes and
    ``check()`` (called by ``close()``, or by the user at any synchronisation point) raises RuntimeError."""

    fuses_backward = True

    def __init__(self, count_max: int, device: torch.device, group=None, average: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (it carries the IPC handles)")
        self.group, self.device = group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("PeerGradComm covers one NVSwitch box (world <= 8)")
        self.scale = 1.0 / self.world if average else 1.0
        self.count_max = (int(count_max) + 3) // 4 * 4
        self.epoch = 0
        self.launches = 0
        L = _lib.lib()
        own = ctypes.c_void_p()
        _lib.check(L.ctvq_peer_alloc(ctypes.byref(own), self.count_max, self.world, device.index), "ctvq_peer_alloc")
        self._own = own
        handle = ctypes.create_string_buffer(64)
        _lib.check(L.ctvq_peer_export(own, handle, device.index), "ctvq_peer_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self._peers = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self._peers.append(own)
            else:
                ptr = ctypes.c_void_p()
                _lib.check(L.ctvq_peer_import(ctypes.create_string_buffer(h, 64), ctypes.byref(ptr), device.index),
                           "ctvq_peer_import")
                self._peers.append(ptr)
        self._table = (ctypes.c_void_p * self.world)(*[p.value for p in self._peers])
        dist.barrier(group=group)  # every rank has mapped every buffer before anyone signals

    def peer_args(self, count: int):
        """-> (table, world, rank, count_max, epoch, scale) for the next collective (advances the epoch)."""
        if count > self.count_max:
            raise RuntimeError(f"codebook gradient ({count} floats) exceeds the symmetric buffer ({self.count_max})")
        self.epoch += 1
        self.launches += 1
        return self._table, self.world, self.rank, self.count_max, self.epoch, self.scale

    def allreduce_(self, grad: torch.Tensor) -> torch.Tensor:
        """Stand-alone form (the local gradient already exists): one small kernel on the current stream."""
        g = grad.detach().contiguous()
        out = torch.empty_like(g, dtype=torch.float32)
        table, world, rank, cmax, epoch, scale = self.peer_args(g.numel())
        sp = _lib.stream_ptr(self.device)
        ws = _lib.workspace(self.device, sp)
        rc = _lib.lib().ctvq_peer_allreduce(table, world, rank, cmax, g.data_ptr(), g.numel(), epoch, scale,
                                            out.data_ptr(), ws.data_ptr(), ws.numel(), self.device.index, sp)
        _lib.check(rc, "ctvq_peer_allreduce")
        return out

    def wait(self) -> None:
        """Kept for callers written against the round-1 interface: the collective is stream-ordered now, nothing to wait for."""

    def check(self) -> None:
        """Synchronises; raises when a peer timed out in any collective since the last check."""
        for (idx, sp), ws in list(_lib._workspaces.items()):
            if idx == self.device.index and _lib.read_and_clear_err(ws, self.device, sp) & 2:
                raise RuntimeError("ct_vae_b200: a peer rank did not reach the codebook-gradient all-reduce within "
                                   "CTVQ_PEER_TIMEOUT_MS; the reduced gradient of that step is incomplete")

    def close(self):
        L = _lib.lib()
        torch.cuda.synchronize(self.device)
        try:
            self.check()
        finally:
            for r, p in enumerate(self._peers):
                if r != self.rank and p is not None:
                    L.ctvq_peer_close(p, self.device.index)
            if self._own is not None:
                L.ctvq_peer_free(self._own, self.device.index)
            self._peers, self._own = [], None


def shard_batch(global_batch: int, world: int, rank: int):
    """Contiguous even split of the batch dimension (DistributedSampler-style, datasets/transition.py:173-176);
    the remainder goes to the lowest ranks.  -> (start, stop)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
