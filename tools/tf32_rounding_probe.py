"""Probe how tcgen05 kind::tf32 reduces fp32 operands to tf32 (truncate vs round-to-nearest): informs the rigorous
error bound of the candidate filter.  Uses the debug dump of the generic tcgen05 kernel."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg
from ct_vae_b200 import _lib

dev = torch.device("cuda:0")
B, D, H, W, K = 2, 32, 8, 8, 64
m = pkg.VectorQuantizerMS(K, D).to(dev)
E = torch.zeros(K, D, device=dev)
E[:, 0] = 1.0
E[1, 0] = 1.0 + 2.0 ** -11 + 2.0 ** -20  # operand B probe
m.embedding.weight.data = E
z = torch.zeros(B, D, H, W, device=dev)
vals = [1.0 + 2.0 ** -11 + 2.0 ** -20, 1.0 + 2.0 ** -11 - 2.0 ** -20, 1.0 + 2.0 ** -10 + 2.0 ** -11, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11]
for i, v in enumerate(vals):
    z[0, 0, 0, i] = v
z[0, 0, 1, 0] = 1.0
Kpad = 64
dump = torch.full((128, Kpad), float("nan"), device=dev)
L = _lib.lib()
L.ctvq_debug_set_tc_dump(ctypes.c_void_p(dump.data_ptr()))
_lib.set_path(_lib.PATH_TC)
m(z, inds=True)
torch.cuda.synchronize()
L.ctvq_debug_set_tc_dump(None)
for i, v in enumerate(vals):
    got = float(dump[i, 0])
    trunc = float(torch.tensor(v).view(torch.int32).bitwise_and(~0x1FFF).view(torch.float32))
    print(f"A operand {v!r:>22}: tc={got!r:>20}  trunc={trunc!r:>20}  rn10={round((v - 1) * 1024) / 1024 + 1!r}")
print(f"B operand {float(E[1,0])!r}: tc={float(dump[8, 1])!r} (z=1.0 row)")
