#!/usr/bin/env python
"""Back-to-back forward-only launches of one shape (for ncu --cache-control none captures).  usage: b2b.py [cfg3|cfg2|cfg1] [iters]"""
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
N, D, K, C, HW, kind = {"cfg3": (1 << 20, 128, 64, 1, 64, "trained"), "cfg2": (1 << 20, 128, 64, 4, 64, "trained"),
                        "cfg1": (1 << 20, 64, 512, 1, 256, "init"), "k16384d256": (1 << 16, 256, 16384, 1, 256, "trained"),
                        "k1024d64": (1 << 20, 64, 1024, 1, 256, "trained"), "k16384d32": (1 << 20, 32, 16384, 1, 256, "trained")}[what]
B = N // HW
side = int(HW ** 0.5)
torch.manual_seed(0)
m = (pkg.MultipleCodebookVectorQuantizer(K, D, C) if C > 1 else pkg.VectorQuantizerMS(K, D)).to(dev)
books = [q.embedding.weight for q in m.quantizers] if C > 1 else [m.embedding.weight]
if kind == "trained":
    for e in books:
        e.data = torch.randn(K, D // C, device=dev) * 0.5
z = torch.randn(B, D, side, side, device=dev)
with torch.no_grad():
    for _ in range(iters):
        m(z, inds=True)
torch.cuda.synchronize()
print("ok")
