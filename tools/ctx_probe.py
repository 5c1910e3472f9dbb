#!/usr/bin/env python
"""Context dependence of a forward-only timing: back-to-back launches (the sustained, power-capped regime: see
tools/b2b_clock.py) vs a launch that follows an unrelated 512 MB write vs a launch that follows the shape's own backward.
usage: ctx_probe.py [cfg3|cfg2|cfg1]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
N, D, K, C, HW, kind = {"cfg3": (1 << 20, 128, 64, 1, 64, "trained"), "cfg2": (1 << 20, 128, 64, 4, 64, "trained"),
                        "cfg1": (1 << 20, 64, 512, 1, 256, "init")}[what]
B = N // HW
side = int(HW ** 0.5)
torch.manual_seed(0)
m = (pkg.MultipleCodebookVectorQuantizer(K, D, C) if C > 1 else pkg.VectorQuantizerMS(K, D)).to(dev)
books = [q.embedding.weight for q in m.quantizers] if C > 1 else [m.embedding.weight]
if kind == "trained":
    for e in books:
        e.data = torch.randn(K, D // C, device=dev) * 0.5
z = torch.randn(B, D, side, side, device=dev)
zg = z.clone().requires_grad_(True)
g = torch.randn_like(z)
one = torch.ones((), device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def run(mode, iters=30):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters + 5):
        a, b = evs[max(0, i - 5)]
        if mode == "flush":
            flush.zero_()
        if mode == "after_bwd":
            o, l = m(zg)
            torch.autograd.backward([o, l], [g, one])
            zg.grad = None
            for e in books:
                e.grad = None
        with torch.no_grad():
            a.record()
            m(z, inds=True)
            b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return round(ts[len(ts) // 2], 4)


print(json.dumps({"what": what, **{mode: run(mode) for mode in ("b2b", "flush", "after_bwd")}}))
