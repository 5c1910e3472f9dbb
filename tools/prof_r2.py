"""Round-2 profiling driver: a few iterations of each workload whose dominant kernels are captured with
`ncu --set full -k regex:...` (see profiles/README.md for the exact command lines).  PROF_WHAT selects the workload:
cfg2 (bench workload, fp32), cfg2_bf16, cfg1 (K=512, D=64), cfg3 (C=1, d=128, K=64)."""
import os
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg  # noqa: E402

what = os.environ.get("PROF_WHAT", "cfg2")
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, D, H, W, C, K, dt, kind = {
    "cfg2": (16384, 128, 8, 8, 4, 64, torch.float32, "trained"),
    "cfg2_bf16": (16384, 128, 8, 8, 4, 64, torch.bfloat16, "trained"),
    "cfg1": (4096, 64, 16, 16, 1, 512, torch.float32, "init"),
    "cfg3": (16384, 128, 8, 8, 1, 64, torch.float32, "trained"),
}[what]
d = D // C
m = (pkg.MultipleCodebookVectorQuantizer(K, D, C) if C > 1 else pkg.VectorQuantizerMS(K, D)).to(dev)
books = [q.embedding.weight for q in m.quantizers] if C > 1 else [m.embedding.weight]
if kind == "trained":
    for e in books:
        e.data = torch.randn(K, d, device=dev) * 0.5
z = torch.randn(B, D, H, W, device=dev).to(dt).requires_grad_(True)
g = torch.randn(B, C * d, H, W, device=dev).to(dt)
one = torch.ones((), device=dev)
for _ in range(4):
    o, l = m(z)
    torch.autograd.backward([o, l], [g, one])
    z.grad = None
    for e in books:
        e.grad = None
torch.cuda.synchronize()
print("ok", what)
