"""fwd / fwd+bwd timing of a few single-codebook shapes (A/B aid for backward-kernel switches such as CTVQ_BWD_C1_GACC)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg  # noqa: E402
from tools.sweep import time_ms  # noqa: E402

dev = torch.device("cuda:0")
for N, D, K, HW, kind in [(1 << 20, 64, 512, 256, "init"), (1 << 20, 128, 64, 64, "trained"), (1 << 20, 32, 256, 256, "trained"),
                          (1 << 20, 64, 256, 256, "trained")]:
    torch.manual_seed(0)
    m = pkg.VectorQuantizerMS(K, D).to(dev)
    if kind == "trained":
        m.embedding.weight.data = torch.randn(K, D, device=dev) * 0.5
    side = int(HW ** 0.5)
    z = torch.randn(N // HW, D, side, side, device=dev, requires_grad=True)
    g = torch.randn(N // HW, D, side, side, device=dev)
    one = torch.ones((), device=dev)

    def fwd():
        with torch.no_grad():
            return m(z)

    def fb():
        o, l = m(z)
        torch.autograd.backward([o, l], [g, one])
        z.grad = None
        m.embedding.weight.grad = None

    tf, tfb = time_ms(fwd, 20, None), time_ms(fb, 20, None)
    bwd_bytes = (3 * D * 4 + 8) * N
    print(json.dumps(dict(N=N, D=D, K=K, fwd_ms=round(tf, 4), fwdbwd_ms=round(tfb, 4), bwd_ms=round(tfb - tf, 4),
                          bwd_frac=round(bwd_bytes / ((tfb - tf) * 1e-3) / 1e9 / 6549.1, 3))), flush=True)
