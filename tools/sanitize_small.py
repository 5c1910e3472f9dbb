"""Tiny run of every kernel (for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py`)."""
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg
from ct_vae_b200 import _lib, gaussian

dev = torch.device("cuda:0")
torch.manual_seed(0)


def fb(m, z, label):
    z = z.clone().requires_grad_(True)
    out, loss, inds = m(z, inds=True)
    (out.sum() * 0.01 + loss).backward()
    torch.cuda.synchronize()
    print(label, "ok", float(loss), "path", _lib.last_path())


for path in (_lib.PATH_AUTO, _lib.PATH_SIMT):
    _lib.set_path(path)
    m = pkg.MultipleCodebookVectorQuantizer(64, 128, 4).to(dev)          # config 2: fast tcgen05 fwd + fast bwd
    fb(m, torch.randn(6, 128, 8, 8, device=dev), "cfg2")
    m = pkg.VectorQuantizerMS(512, 64).to(dev)                            # config 1: c1 kernel (two rounds)
    fb(m, torch.randn(3, 64, 16, 16, device=dev), "cfg1")
    m = pkg.MultipleCodebookVectorQuantizer(64, 128, 1).to(dev)          # config 3
    fb(m, torch.randn(5, 128, 8, 8, device=dev), "cfg3")
    m = pkg.MultipleCodebookVectorQuantizer(50, 48, 2).to(dev)           # generic tcgen05 kernel (K padded, d=24)
    m.chan_stride = 24
    fb(m, torch.randn(3, 48, 8, 4, device=dev), "generic")
    m = pkg.MultipleCodebookVectorQuantizer(7, 15, 5).to(dev)            # ragged: SIMT + generic backward
    fb(m, torch.randn(2, 15, 3, 3, device=dev), "ragged")
_lib.set_path(_lib.PATH_AUTO)
m = pkg.MultipleCodebookVectorQuantizer(64, 128, 4).to(dev)
z = torch.randn(4, 128, 8, 8, device=dev)
i = m.compute_inds(z)
ix, iy = m.compute_inds_pair(z, z * 1.5)
q, l = m.compute_latents(z, i)
mu, lv = torch.randn(8, 16, device=dev, requires_grad=True), torch.randn(8, 16, device=dev, requires_grad=True)
zz, k = gaussian.reparam_kld(mu, lv)
(zz.sum() + k).backward()
torch.cuda.synchronize()
print("all ok")
