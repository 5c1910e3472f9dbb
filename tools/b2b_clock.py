#!/usr/bin/env python
"""Per-iteration forward-only times of back-to-back launches with nvidia-smi clock / power samples taken alongside."""
import json
import subprocess
import sys
import time

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
N, D, K, C, HW, kind = {"cfg3": (1 << 20, 128, 64, 1, 64, "trained"), "cfg2": (1 << 20, 128, 64, 4, 64, "trained")}[what]
B = N // HW
torch.manual_seed(0)
m = (pkg.MultipleCodebookVectorQuantizer(K, D, C) if C > 1 else pkg.VectorQuantizerMS(K, D)).to(dev)
books = [q.embedding.weight for q in m.quantizers] if C > 1 else [m.embedding.weight]
for e in books:
    e.data = torch.randn(K, D // C, device=dev) * 0.5
z = torch.randn(B, D, 8, 8, device=dev)
with torch.no_grad():
    m(z, inds=True)
torch.cuda.synchronize()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "20"],
                       stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
iters = 3000
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
with torch.no_grad():
    for a, b in evs:
        a.record()
        m(z, inds=True)
        b.record()
torch.cuda.synchronize()
time.sleep(0.1)
smi.terminate()
lines = smi.stdout.read().strip().splitlines()
ts = [a.elapsed_time(b) for a, b in evs]
print(json.dumps({"what": what, "first10": [round(t, 4) for t in ts[:10]], "i100": round(ts[100], 4), "i500": round(ts[500], 4),
                  "i1500": round(ts[1500], 4), "last": round(ts[-1], 4), "median": round(sorted(ts)[iters // 2], 4), "min": round(min(ts), 4)}))
print("smi:", " | ".join(lines[::4][:16]))
