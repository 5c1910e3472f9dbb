"""Profiling driver of the resident-codebook single-codebook forward (ctvq_tc_res.cu) at the config-1 codebook
(K=512, D=64, init-scale), 1 M latents: `ncu --set full -k regex:vq_fwd_tc_res -s 2 -c 1 python tools/prof_res.py`."""
import os
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
K, D = int(os.environ.get("PROF_K", 512)), int(os.environ.get("PROF_D", 64))
m = pkg.VectorQuantizerMS(K, D).to(dev)
if os.environ.get("PROF_TRAINED"):
    m.embedding.weight.data = torch.randn(K, D, device=dev) * 0.5
z = torch.randn(4096, D, 16, 16, device=dev)
with torch.no_grad():
    for _ in range(4):
        m(z, inds=True)
torch.cuda.synchronize()
