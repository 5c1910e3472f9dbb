#!/usr/bin/env python
"""Config-4 microbenchmark sweep (BASELINE.json configs[3]) + config-5 Gaussian branch, on one B200.

For each (N, D, K, C) point: forward (fused argmin+gather+loss), forward+backward, the kernel path taken, achieved
algorithmic GB/s and the fraction of the measured HBM peak, index parity vs the C oracle on a 4096-row sample.
Writes a markdown table (default profiles/r1_sweep.md).  CUDA-event timing, 3 warm-ups, inputs larger than L2 or
L2 flushed between iterations.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ct_vae_b200 as pkg  # noqa: E402
from ct_vae_b200 import _lib, ct_codec, gaussian  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402

PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_ms(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        if flush is not None:
            flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def point(N, D, K, C, HW, kind, dev, flush, dtype=torch.float32):
    d = D // C
    B = N // HW
    torch.manual_seed(0)
    m = (pkg.MultipleCodebookVectorQuantizer(K, D, C) if C > 1 else pkg.VectorQuantizerMS(K, D)).to(dev)
    books = [q.embedding.weight for q in m.quantizers] if C > 1 else [m.embedding.weight]
    if kind == "trained":
        for e in books:
            e.data = torch.randn(K, d, device=dev) * 0.5
    side = int(HW ** 0.5)
    z = torch.randn(B, D, side, side, device=dev).to(dtype).requires_grad_(True)
    g_out = torch.randn(B, C * d, side, side, device=dev).to(dtype)
    g_loss = torch.ones((), device=dev)
    used = min(D, (C - 1) + d)
    es = 2 if dtype == torch.bfloat16 else 4   # bytes per latent / output / gradient element (indices stay int64)
    fb = es * used + es * C * d + 8 * C
    bb = es * C * d + es * used + 8 * C + es * D
    small = N * D * es < 200e6
    fl = flush if small else None

    def fwd():
        with torch.no_grad():
            return m(z, inds=True)

    def fwdbwd():
        out, loss = m(z)
        torch.autograd.backward([out, loss], [g_out, g_loss])
        z.grad = None
        for e in books:
            e.grad = None

    iters = 10 if N >= (1 << 22) else 20
    t_f = time_ms(fwd, iters, fl)
    path = {1: "simt", 2: "tcgen05"}.get(_lib.last_path(), "?")
    t_fb = time_ms(fwdbwd, iters, fl)
    # parity sample vs the C oracle (exact)
    with torch.no_grad():
        _, _, inds = m(z[: max(1, 4096 // HW)], inds=True)
    ref = CO.argmin(z[: max(1, 4096 // HW)].detach().float().cpu(), [e.detach().to(dtype).float().cpu() for e in books])
    ok = bool(torch.equal(inds.cpu().reshape(ref.shape), ref))
    return dict(N=N, D=D, K=K, C=C, HW=HW, codebook=kind, dtype="bf16" if dtype == torch.bfloat16 else "fp32", path=path, fwd_ms=t_f, fwdbwd_ms=t_fb,
                fwd_Mlat_s=N / t_f / 1e3, fwdbwd_Mlat_s=N / t_fb / 1e3, fwd_gbs=fb * N / t_f / 1e6,
                fwd_frac=fb * N / t_f / 1e6 / PEAK, fwd_tflops=2.0 * N * K * d * C / t_f / 1e9, fwdbwd_gbs=(fb + bb) * N / t_fb / 1e6,
                fwdbwd_frac=(fb + bb) * N / t_fb / 1e6 / PEAK, idx_exact=ok)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_sweep.md"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    pts = [
        # (N, D, K, C, HW, codebook)   configs 1-3 at their own batch and at a bandwidth-meaningful batch
        (64 * 256, 64, 512, 1, 256, "init"), (1 << 20, 64, 512, 1, 256, "init"),           # configs/vq_vae.yaml
        (64 * 64, 128, 64, 4, 64, "trained"), (1 << 20, 128, 64, 4, 64, "trained"), (1 << 22, 128, 64, 4, 64, "trained"),  # mcq_vae.yaml
        (16 * 64, 128, 64, 1, 64, "trained"), (1 << 20, 128, 64, 1, 64, "trained"),       # ct_mcq_vae.yaml
        # config-4 sweep
        (1 << 16, 32, 256, 1, 256, "trained"), (1 << 20, 32, 256, 1, 256, "trained"), (1 << 22, 32, 256, 1, 256, "trained"),
        (1 << 20, 64, 256, 1, 256, "trained"), (1 << 20, 128, 256, 1, 256, "trained"), (1 << 18, 256, 256, 1, 256, "trained"),
        (1 << 20, 64, 1024, 1, 256, "trained"), (1 << 18, 128, 4096, 1, 256, "trained"), (1 << 16, 256, 16384, 1, 256, "trained"),
        (1 << 20, 32, 16384, 1, 256, "trained"),
        # neighbours of the MCQ shape (two codebooks; 128x128 images) and bf16 latents (dtype = CTVQ_BF16)
        (1 << 20, 64, 64, 2, 64, "trained"), (1 << 20, 128, 64, 4, 256, "trained"),
        (1 << 20, 128, 64, 4, 64, "trained", torch.bfloat16), (1 << 20, 128, 64, 1, 64, "trained", torch.bfloat16),
        (1 << 20, 64, 512, 1, 256, "init", torch.bfloat16),
    ]
    if not args.quick:
        pts += [(1 << 24, 32, 256, 1, 256, "trained")]
    rows = []
    for pt in pts:
        try:
            r = point(*pt[:6], dev, flush, *(pt[6:]))
        except RuntimeError as e:
            r = dict(N=pt[0], D=pt[1], K=pt[2], C=pt[3], HW=pt[4], codebook=pt[5], dtype="?", error=str(e)[:80])
        rows.append(r)
        print(json.dumps(r), flush=True)
    # config 5: reparam + KL
    g = []
    for B, L in ((4096, 128), (1 << 17, 128)):
        mu = torch.randn(B, L, device=dev, requires_grad=True)
        lv = (torch.randn(B, L, device=dev) * 0.5).requires_grad_(True)
        eps = torch.randn(B, L, device=dev)
        gz = torch.randn(B, L, device=dev)
        gk = torch.ones((), device=dev)
        fl = flush if B * L * 16 < 200e6 else None

        def f():
            with torch.no_grad():
                return gaussian.reparam_kld(mu, lv, eps)

        def fb():
            zz, kk = gaussian.reparam_kld(mu, lv, eps)
            torch.autograd.backward([zz, kk], [gz, gk])
            mu.grad = None
            lv.grad = None

        tf_, tfb = time_ms(f, 20, fl), time_ms(fb, 20, fl)
        r = dict(B=B, L=L, fwd_ms=tf_, fwdbwd_ms=tfb, fwd_gbs=16 * B * L / tf_ / 1e6, fwd_frac=16 * B * L / tf_ / 1e6 / PEAK,
                 fwdbwd_gbs=(16 + 24) * B * L / tfb / 1e6, fwdbwd_frac=(16 + 24) * B * L / tfb / 1e6 / PEAK)
        g.append(r)
        print(json.dumps(r), flush=True)
    # CT-mode codec: one-hot <-> index converters and the one-hot cross-entropy
    ct = []
    for B, C, H, W, N in ((16, 1, 8, 8, 64), (16384, 4, 8, 8, 64)):
        shape = [B, 128, H, W]
        inds = torch.randint(0, N, (B, C, H, W), device=dev)
        onehot = ct_codec.ct_preprocess(inds, shape, N, C)
        lat = (torch.rand(B, N, C * H, W, device=dev) * 0.1).requires_grad_(True)
        lat_y = torch.rand(B, N, C * H, W, device=dev)
        one = torch.ones((), device=dev)
        nrows = B * C * H * W
        fl = flush if nrows * N * 4 < 200e6 else None

        def ce_fb():
            loss = ct_codec.latent_cross_entropy_loss(lat, lat_y)
            torch.autograd.backward([loss], [one])
            lat.grad = None

        t_pre = time_ms(lambda: ct_codec.ct_preprocess(inds, shape, N, C), 20, fl)
        t_post = time_ms(lambda: ct_codec.ct_postprocess(onehot, shape, N, C), 20, fl)
        t_ce = time_ms(ce_fb, 20, fl)
        r = dict(B=B, C=C, HW=H * W, N=N, rows=nrows, pre_ms=t_pre, pre_gbs=nrows * (4 * N + 8) / t_pre / 1e6,
                 post_ms=t_post, post_gbs=nrows * (4 * N + 8) / t_post / 1e6, ce_ms=t_ce,
                 ce_gbs=nrows * (8 * N + 12 + 4 * N + 12 + 4 * N) / t_ce / 1e6)
        ct.append(r)
        print(json.dumps(r), flush=True)
    with open(args.out, "w") as f:
        f.write("# round 2 — quantiser microbenchmark sweep (BASELINE.json configs[3]) and Gaussian branch (configs[4])\n\n")
        f.write(f"One B200, fp32, CUDA-event median, HBM peak {PEAK:.0f} GB/s (MEASURED_PEAKS.json). `frac` = algorithmic bytes ÷ time ÷ peak "
                "(HBM roofline). `fwd TF/s` = 2*K*D flops per row / time: the points with K >= 1024 are TENSOR-bound (tf32 peak taken as "
                "half the measured bf16 burst peak of 1670 TF/s, i.e. 835 TF/s) and run the streaming tcgen05 kernel. "
                "`idx_exact` = indices equal to the C oracle on a 4096-row sample.  bf16 rows: latents / outputs / gradients are bf16 (dtype = CTVQ_BF16), "
                "the algorithmic bytes are counted at 2 bytes per element.\n\n")
        f.write("| N | D | K | C | HW | codebook | dtype | path | fwd ms | fwd M lat/s | fwd GB/s | fwd frac | fwd TF/s | fwd+bwd ms | fwd+bwd M lat/s | fwd+bwd GB/s | fwd+bwd frac | idx_exact |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            if "error" in r:
                f.write(f"| {r['N']} | {r['D']} | {r['K']} | {r['C']} | {r['HW']} | {r['codebook']} | error: {r['error']} |\n")
                continue
            f.write(f"| {r['N']} | {r['D']} | {r['K']} | {r['C']} | {r['HW']} | {r['codebook']} | {r['dtype']} | {r['path']} | {r['fwd_ms']:.4f} | "
                    f"{r['fwd_Mlat_s']:.1f} | {r['fwd_gbs']:.0f} | {r['fwd_frac']:.3f} | {r['fwd_tflops']:.0f} | {r['fwdbwd_ms']:.4f} | {r['fwdbwd_Mlat_s']:.1f} | "
                    f"{r['fwdbwd_gbs']:.0f} | {r['fwdbwd_frac']:.3f} | {r['idx_exact']} |\n")
        f.write("\n## reparameterise + KL (16 B/element forward: mu, logvar, eps in, z out; backward +24 B)\n\n")
        f.write("| B | L | fwd ms | fwd GB/s | fwd frac | fwd+bwd ms | fwd+bwd GB/s | fwd+bwd frac |\n|---|---|---|---|---|---|---|---|\n")
        for r in g:
            f.write(f"| {r['B']} | {r['L']} | {r['fwd_ms']:.4f} | {r['fwd_gbs']:.0f} | {r['fwd_frac']:.3f} | {r['fwdbwd_ms']:.4f} | "
                    f"{r['fwdbwd_gbs']:.0f} | {r['fwdbwd_frac']:.3f} |\n")
        f.write("\n## CT-mode codec (models/ct_mcq_vae.py:472-496, 306-311): one-hot [B,N,C*H,W] fp32 <-> indices, one-hot cross-entropy fwd+bwd\n\n")
        f.write("| B | C | HW | N | rows | ct_preprocess ms | GB/s | frac | ct_postprocess ms | GB/s | frac | latent CE fwd+bwd ms | GB/s | frac |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in ct:
            f.write(f"| {r['B']} | {r['C']} | {r['HW']} | {r['N']} | {r['rows']} | {r['pre_ms']:.4f} | {r['pre_gbs']:.0f} | {r['pre_gbs'] / PEAK:.3f} | "
                    f"{r['post_ms']:.4f} | {r['post_gbs']:.0f} | {r['post_gbs'] / PEAK:.3f} | {r['ce_ms']:.4f} | {r['ce_gbs']:.0f} | {r['ce_gbs'] / PEAK:.3f} |\n")
    print("wrote", args.out)


if __name__ == "__main__":
    main()
