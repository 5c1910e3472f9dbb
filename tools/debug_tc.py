"""Debug aid: dump the raw tcgen05 accumulators of the TC forward kernel and compare with z @ E^T."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg
from ct_vae_b200 import _lib

B, D, H, W, C, K = [int(x) for x in (sys.argv[1:7] if len(sys.argv) > 6 else (4, 32, 8, 8, 1, 64))]
d = D // C
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = pkg.MultipleCodebookVectorQuantizer(K, D, C).to(dev)
for q in m.quantizers:
    q.embedding.weight.data = torch.randn(K, d, device=dev) * 0.5
z = torch.randn(B, D, H, W, device=dev)
Kpad = (K + 15) // 16 * 16
ntiles = (B * H * W + 127) // 128
dump = torch.full((ntiles * 128, C * Kpad), float("nan"), device=dev)
L = _lib.lib()
L.ctvq_debug_set_tc_dump(ctypes.c_void_p(dump.data_ptr()))
_lib.set_path(_lib.PATH_TC)
out, loss, inds = m(z, inds=True)
torch.cuda.synchronize()
L.ctvq_debug_set_tc_dump(None)
rows = z.permute(0, 2, 3, 1).reshape(-1, D)
for c, q in enumerate(m.quantizers):
    ref = rows[:, c:c + d] @ q.embedding.weight.t()
    got = dump[: rows.shape[0], c * Kpad: c * Kpad + K]
    err = (got - ref).abs()
    print(f"codebook {c}: max |dot_tc - dot_ref| = {float(err.max()):.4e}  (ref max {float(ref.abs().max()):.3f}); nan={int(torch.isnan(got).sum())}")
    if float(err.max()) > 0.05:
        print(" got[0,:8]", got[0, :8].tolist())
        print(" ref[0,:8]", ref[0, :8].tolist())
        # try to recognise permutations: does got row r match ref row r' ?
        g0 = got[0]
        best = ((ref - g0).abs().sum(1)).argmin()
        print(" got row 0 is closest to ref row", int(best), "err", float((ref[best] - g0).abs().max()))
        # column check
        gc = got[:, 0]
        bc = ((ref - gc[:, None]).abs().sum(0)).argmin()
        print(" got col 0 is closest to ref col", int(bc), "err", float((ref[:, bc] - gc).abs().max()))
from oracle import c_oracle as CO
ci = CO.argmin(z.cpu(), [q.embedding.weight.detach().cpu() for q in m.quantizers])
print("index mismatches vs C oracle:", int((ci != inds.cpu()).sum()), "of", ci.numel())
print("dump: frac zero =", float((dump[: rows.shape[0]] == 0).float().mean()), " frac nan =", float(torch.isnan(dump[: rows.shape[0]]).float().mean()))
books = [q.embedding.weight.detach() for q in m.quantizers]
exp = torch.cat([(rows[:, c:c + d] + (books[c][inds[:, c].reshape(-1)] - rows[:, c:c + d])) for c in range(C)], 1)
got_out = out.detach().permute(0, 2, 3, 1).reshape(-1, C * d)
print("out vs z+(E[idx]-z) at the kernel's own indices: max err", float((got_out - exp).abs().max()))
