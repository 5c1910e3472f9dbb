"""ctypes binding of libctvq.so (include/ctvq.h).  There is NO fallback: if the library is missing or a
tensor is not on a CUDA device the ops raise."""
import ctypes
import os
import threading
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libctvq.so")
_lib = None
_lock = threading.Lock()

PATH_AUTO, PATH_SIMT, PATH_TC, PATH_TC_STREAM = 0, 1, 2, 3
F32, BF16 = 0, 1

_vp, _i, _i64, _sz, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float
_SIGNATURES = {
    "ctvq_version": (_i, []),
    "ctvq_strerror": (ctypes.c_char_p, [_i]),
    "ctvq_workspace_bytes": (_sz, [_i, _i, _i]),
    "ctvq_set_path": (_i, [_i]),
    "ctvq_last_path": (_i, []),
    "ctvq_read_and_clear_err": (_i, [_vp, _sz, ctypes.POINTER(ctypes.c_uint), _i, _vp]),
    "ctvq_argmin": (_i, [_vp, _i, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "ctvq_gather_st_loss": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _i, _vp]),
    "ctvq_forward": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "ctvq_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _i, _vp]),
    "ctvq_reparam_kld_fwd": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "ctvq_reparam_kld_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _i, _vp]),
    "ctvq_onehot_from_inds": (_i, [_vp, _i64, _i64, _i, _vp, _vp, _sz, _i, _vp]),
    "ctvq_inds_from_onehot": (_i, [_vp, _i64, _i64, _i, _vp, _i, _vp]),
    "ctvq_latent_ce_fwd": (_i, [_vp, _vp, _i64, _i64, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "ctvq_latent_ce_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _i, _vp, _i, _vp]),
    "ctvq_nccl_load": (_i, [ctypes.c_char_p]),
    "ctvq_nccl_unique_id": (_i, [_vp]),
    "ctvq_nccl_comm_init": (_i, [ctypes.POINTER(_vp), _i, _i, _vp, _i]),
    "ctvq_nccl_comm_destroy": (_i, [_vp]),
    "ctvq_allreduce_codebook_grad": (_i, [_vp, _vp, _sz, _f, _i, _vp]),
    "ctvq_debug_set_fast_trace": (None, [_vp]),
    "ctvq_debug_set_tc_dump": (None, [_vp]),
    "ctvq_peer_buffer_bytes": (_sz, [_sz, _i]),
    "ctvq_peer_alloc": (_i, [ctypes.POINTER(_vp), _sz, _i, _i]),
    "ctvq_peer_free": (_i, [_vp, _i]),
    "ctvq_peer_export": (_i, [_vp, _vp, _i]),
    "ctvq_peer_import": (_i, [_vp, ctypes.POINTER(_vp), _i]),
    "ctvq_peer_close": (_i, [_vp, _i]),
    "ctvq_peer_allreduce": (_i, [_vp, _i, _i, _sz, _vp, _sz, ctypes.c_uint, _f, _vp, _vp, _sz, _i, _vp]),
    "ctvq_backward_allreduce": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp,
                                     _vp, _i, _i, _sz, ctypes.c_uint, _f, _vp, _vp, _sz, _i, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load libctvq.so (built by ``__graft_entry__.build()`` / ``make -C ct_vae_b200/csrc``)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"ct_vae_b200: {LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (there is no CPU or PyTorch fallback for the quantiser path)")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int, what: str = "ctvq"):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {lib().ctvq_strerror(rc).decode()}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ct_vae_b200 ops run on CUDA tensors only — there is no CPU fallback "
                               f"(got a tensor on {t.device})")


_workspaces = {}


def workspace(device: torch.device, stream_ptr: int, c: int = 0, k: int = 0, d: int = 0) -> torch.Tensor:
    """Zero-initialised per-(device, stream) scratch the kernels keep self-cleaning; grown (never shrunk) to what
    ctvq_workspace_bytes asks for the problem at hand (the streaming single-codebook kernel keeps |e|^2 there)."""
    key = (device.index, stream_ptr)
    ws = _workspaces.get(key)
    n = lib().ctvq_workspace_bytes(c, k, d)
    if ws is None or ws.numel() < n:
        ws = torch.zeros(max(n, 4096), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_validate = os.environ.get("CTVQ_VALIDATE", "0") not in ("", "0")


def set_validate(on: bool) -> bool:
    """Debug switch (also env CTVQ_VALIDATE=1): after every op that consumes caller-supplied indices, read the
    error word back (one stream synchronisation per call) and raise IndexError like the reference's ``scatter_`` /
    ``F.one_hot`` do (models/vq_vae.py:40, models/ct_mcq_vae.py:480).  Returns the previous setting."""
    global _validate
    prev, _validate = _validate, bool(on)
    return prev


def validating() -> bool:
    return _validate


def read_and_clear_err(ws: torch.Tensor, device: torch.device, sp: int) -> int:
    word = ctypes.c_uint(0)
    check(lib().ctvq_read_and_clear_err(ws.data_ptr(), ws.numel(), ctypes.byref(word), device.index, sp),
          "ctvq_read_and_clear_err")
    return int(word.value)


def raise_if_bad_indices(device: Optional[torch.device] = None) -> None:
    """Explicit check point (synchronises): raises IndexError when any op on ``device`` (default: every device) met an
    index outside [0, K) since the last check.  The kernels clamp such indices instead of faulting."""
    bad = []
    for (idx, sp), ws in list(_workspaces.items()):
        if device is not None and device.index != idx:
            continue
        if read_and_clear_err(ws, torch.device("cuda", idx), sp):
            bad.append(idx)
    if bad:
        raise IndexError(f"ct_vae_b200: code index out of range [0, K) on cuda device(s) {sorted(set(bad))} "
                         "(the reference raises from scatter_/F.one_hot; the kernels clamped it)")


def maybe_validate(ws: torch.Tensor, device: torch.device, sp: int, what: str) -> None:
    if _validate and read_and_clear_err(ws, device, sp):
        raise IndexError(f"{what}: code index out of range [0, K) (clamped by the kernel; the reference raises from "
                         "scatter_/F.one_hot)")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def set_path(path: int) -> int:
    return lib().ctvq_set_path(path)


def last_path() -> int:
    return lib().ctvq_last_path()
