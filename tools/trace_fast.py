#!/usr/bin/env python
"""Where does the fixed cost of the config-2 forward kernel go?  Per-CTA %globaltimer stamps (debug hook
ctvq_debug_set_fast_trace): 0 kernel entry, 1 barriers+TMEM ready, 2 codebooks staged, 3 first slab landed,
4 warp 0's first unit done, 5 warp 0's last unit done, 6 CTA exit.  Prints min / median / max over CTAs in us relative
to the earliest entry."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg  # noqa: E402
from ct_vae_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
m = pkg.MultipleCodebookVectorQuantizer(64, 128, 4).to(dev)
z = torch.randn(B, 128, 8, 8, device=dev)
L = _lib.lib()
L.ctvq_debug_set_fast_trace.argtypes = [ctypes.c_void_p]
with torch.no_grad():
    for _ in range(3):
        m(z, inds=True)
    torch.cuda.synchronize()
    tr = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    L.ctvq_debug_set_fast_trace(ctypes.c_void_p(tr.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m(z, inds=True)
    e1.record()
    torch.cuda.synchronize()
    L.ctvq_debug_set_fast_trace(None)
t = tr.cpu().view(148, 8).double()
t0 = t[:, 0].min()
names = ["entry", "barriers+TMEM", "codebooks staged", "first slab landed", "first unit done", "last unit done", "exit"]
print(f"B={B}: event time {e0.elapsed_time(e1) * 1e3:.1f} us (traced launch)")
for i, n in enumerate(names):
    c = (t[:, i] - t0) / 1e3
    print(f"  {n:20s} min {c.min():8.2f}  median {c.median():8.2f}  max {c.max():8.2f} us")
