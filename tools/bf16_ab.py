"""fp32 vs native bf16 (dtype = CTVQ_BF16) on the config-2 shape at 1 M latent rows: forward / forward+backward time,
algorithmic GB/s with the bytes of the respective element size, exact-index check on a sample against the C oracle."""
import json
import sys

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg  # noqa: E402
from ct_vae_b200 import _lib  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402
from tools.sweep import time_ms  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6549.1
for name, B, D, H, W, C, K in [("cfg2", 16384, 128, 8, 8, 4, 64), ("cfg2_hw256", 4096, 128, 16, 16, 4, 64), ("cfg1", 4096, 64, 16, 16, 1, 512)]:
    d = D // C
    for dt in (torch.float32, torch.bfloat16):
        torch.manual_seed(0)
        m = (pkg.MultipleCodebookVectorQuantizer(K, D, C) if C > 1 else pkg.VectorQuantizerMS(K, D)).to(dev)
        books = [q.embedding.weight for q in m.quantizers] if C > 1 else [m.embedding.weight]
        for e in books:
            e.data = torch.randn(K, d, device=dev) * 0.5
        z = torch.randn(B, D, H, W, device=dev).to(dt).requires_grad_(True)
        g = torch.randn(B, C * d, H, W, device=dev).to(dt)
        one = torch.ones((), device=dev)

        def fwd():
            with torch.no_grad():
                return m(z, inds=True)

        def fb():
            o, l = m(z)
            torch.autograd.backward([o, l], [g, one])
            z.grad = None
            for e in books:
                e.grad = None

        tf, tfb = time_ms(fwd, 20, None), time_ms(fb, 20, None)
        path = _lib.last_path()
        with torch.no_grad():
            _, _, inds = m(z[:32], inds=True)
        zr = z[:32].detach().float().cpu()
        er = [(e.detach().to(dt).float().cpu()) for e in books]
        ok = bool(torch.equal(inds.cpu().reshape(32, C, H, W), CO.argmin(zr, er)))
        es = 4 if dt == torch.float32 else 2
        used = min(D, (C - 1) + d)
        fbytes = es * used + es * C * d + 8 * C
        bbytes = es * C * d + es * used + 8 * C + es * D
        N = B * H * W
        print(json.dumps(dict(name=name, dtype=str(dt).split(".")[1], path=path, fwd_ms=round(tf, 4), fwdbwd_ms=round(tfb, 4),
                              fwd_frac=round(fbytes * N / tf / 1e6 / PEAK, 3), fwdbwd_frac=round((fbytes + bbytes) * N / tfb / 1e6 / PEAK, 3),
                              Mlat_s_fwdbwd=round(N / tfb / 1e3), idx_exact=ok, near=m.near_tie_rows())), flush=True)
