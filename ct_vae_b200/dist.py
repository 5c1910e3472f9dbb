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


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias memory the C library allocated."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerGradComm:
    """The same collective WITHOUT NCCL: a one-shot all-reduce over NVLink peer memory (ctvq_peer_* in include/ctvq.h).

    The backward kernel writes grad_E straight into this rank's symmetric slot (`grad_buffer`); `allreduce_` launches
    ONE kernel that handshakes through flag words in the peers' buffers and sums the `world` slots in rank order
    (bit-identical on every rank), scaled by 1/world.  torch.distributed only carries the 64-byte IPC handles.
    Single node, world <= 8 (every GPU reaches every peer through NVSwitch)."""

    def __init__(self, count_max: int, device: torch.device, group=None, average: bool = True, overlap: bool = False):
        """overlap=True launches the all-reduce kernel on a private side stream (ordered after the backward kernel by an
        event), so it runs underneath whatever the caller enqueues next -- the next forward, the encoder's backward.
        The caller must then call ``wait()`` before it READS the reduced gradient (optimizer step); the next backward
        waits by itself before it reuses a slot."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (it carries the IPC handles)")
        self.group, self.device = group, device
        self.overlap = bool(overlap)
        self._side = torch.cuda.Stream(device=device) if overlap else None
        self._ready = torch.cuda.Event() if overlap else None   # backward kernel finished writing the slot
        self._done = None                                        # last all-reduce finished (side stream)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("PeerGradComm covers one NVSwitch box (world <= 8)")
        self.scale = 1.0 / self.world if average else 1.0
        self.count_max = int(count_max)
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

    def grad_buffer(self, shape) -> torch.Tensor:
        """Tensor aliasing the slot the NEXT all-reduce will read (the backward kernel's gE_out)."""
        n = 1
        for s in shape:
            n *= int(s)
        if n > self.count_max:
            raise RuntimeError(f"codebook gradient ({n} floats) exceeds the symmetric buffer ({self.count_max})")
        if self._done is not None:
            # the slot about to be rewritten was last read by the peers during all-reduce (epoch - 1); our all-reduce of
            # `epoch` completing proves every peer has finished that one
            torch.cuda.current_stream(self.device).wait_event(self._done)
        slot = _lib.lib().ctvq_peer_slot(self._own, self.count_max, self.epoch + 1)
        return torch.as_tensor(_DeviceArray(slot, n), device=self.device).view(*shape)

    def allreduce_(self, grad: torch.Tensor) -> torch.Tensor:
        self.epoch += 1
        out = torch.empty(grad.shape, dtype=torch.float32, device=self.device)
        if self.overlap:
            main = torch.cuda.current_stream(self.device)
            self._ready.record(main)
            self._side.wait_event(self._ready)
            sp = self._side.cuda_stream
            out.record_stream(self._side)
        else:
            sp = _lib.stream_ptr(self.device)
        rc = _lib.lib().ctvq_peer_allreduce(self._table, self.world, self.rank, self.count_max, grad.numel(),
                                            self.epoch, self.scale, out.data_ptr(), self.device.index, sp)
        _lib.check(rc, "ctvq_peer_allreduce")
        if self.overlap:
            self._done = torch.cuda.Event()
            self._done.record(self._side)
        self.launches += 1
        return out

    def wait(self) -> None:
        """Make the current stream wait for the last all-reduce (no-op without overlap)."""
        if self._done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done)

    def close(self):
        L = _lib.lib()
        torch.cuda.synchronize(self.device)
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
