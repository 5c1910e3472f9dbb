#!/usr/bin/env python
"""Forward-only timing of single-codebook shapes (run once as is, once with CTVQ_NO_STREAM=1 to compare the streaming
tcgen05 kernel with the older resident-codebook kernels / SIMT fallback).  Prints ms, path and index parity vs the C oracle."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ct_vae_b200 as pkg  # noqa: E402
from ct_vae_b200 import _lib  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402
from tools.sweep import time_ms  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
pts = [(1 << 14, 64, 512, 256, "init"), (1 << 20, 64, 512, 256, "init"), (1 << 20, 64, 512, 256, "trained"),
       (1 << 20, 128, 64, 64, "trained"), (1 << 20, 32, 256, 256, "trained"), (1 << 20, 64, 256, 256, "trained"),
       (1 << 20, 128, 256, 256, "trained"), (1 << 18, 256, 256, 256, "trained"), (1 << 20, 64, 1024, 256, "trained"),
       (1 << 18, 128, 4096, 256, "trained"), (1 << 16, 256, 16384, 256, "trained"), (1 << 20, 32, 16384, 256, "trained")]
for N, D, K, HW, kind in pts:
    torch.manual_seed(0)
    m = pkg.VectorQuantizerMS(K, D).to(dev)
    if kind == "trained":
        m.embedding.weight.data = torch.randn(K, D, device=dev) * 0.5
    side = int(HW ** 0.5)
    z = torch.randn(N // HW, D, side, side, device=dev)

    def fwd():
        with torch.no_grad():
            return m(z, inds=True)

    if os.environ.get("CTVQ_FORCE_STREAM"):
        _lib.set_path(_lib.PATH_TC_STREAM)
    if os.environ.get("CTVQ_ONLY_D") and int(os.environ["CTVQ_ONLY_D"]) != D:
        continue
    t = time_ms(fwd, 10, flush if N * D * 4 < 200e6 else None)
    nb = max(1, 2048 // HW)
    if os.environ.get("CTVQ_FORCE_STREAM"):
        _lib.set_path(_lib.PATH_TC_STREAM)
    with torch.no_grad():
        _, _, inds = m(z[:nb], inds=True)
    ref = CO.argmin(z[:nb].cpu(), [m.embedding.weight.detach().cpu()])
    ok = bool(torch.equal(inds.cpu().reshape(ref.shape), ref))
    tf = 2.0 * N * K * D / t / 1e9
    print(json.dumps(dict(N=N, D=D, K=K, kind=kind, ms=round(t, 4), tflops=round(tf, 1), gbs=round((8 * D + 8) * N / t / 1e6),
                          path=_lib.last_path(), idx_exact=ok, stream=os.environ.get("CTVQ_NO_STREAM") is None)), flush=True)
