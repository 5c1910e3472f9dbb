import sys, torch
sys.path.insert(0, ".")
import ct_vae_b200 as pkg
from ct_vae_b200 import _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = pkg.VectorQuantizerMS(512, 64).to(dev)
z = torch.randn(4096, 64, 16, 16, device=dev)
_lib.set_path(_lib.PATH_TC_STREAM)
with torch.no_grad():
    for _ in range(4):
        m(z, inds=True)
torch.cuda.synchronize()
