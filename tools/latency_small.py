"""Latency of the quantiser at the configs' OWN batch sizes (launch-bound regime): eager call vs CUDA-graph replay."""
import sys
import time

import torch

sys.path.insert(0, ".")
import ct_vae_b200 as pkg

dev = torch.device("cuda:0")


def run(name, B, D, H, W, C, K):
    d = D // C
    m = (pkg.MultipleCodebookVectorQuantizer(K, D, C) if C > 1 else pkg.VectorQuantizerMS(K, D)).to(dev)
    z = torch.randn(B, D, H, W, device=dev, requires_grad=True)
    g = torch.randn(B, C * d, H, W, device=dev)
    gl = torch.ones((), device=dev)
    params = list(m.parameters())

    def fb():
        o, l = m(z)
        torch.autograd.backward([o, l], [g, gl])

    def clear():
        z.grad = None
        for p in params:
            p.grad = None

    for _ in range(5):
        fb(); clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fb(); clear()
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) / 200 * 1e6
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fb(); clear()
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fb()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: N={B*H*W} rows  eager fwd+bwd {eager:.1f} us/call   CUDA-graph replay {a.elapsed_time(b)/200*1e3:.1f} us")


run("config 1 (VQ-VAE, B=64)", 64, 64, 16, 16, 1, 512)
run("config 2 (MCQ-VAE, B=64)", 64, 128, 8, 8, 4, 64)
run("config 3 (CT-MCQ-VAE, B=16)", 16, 128, 8, 8, 1, 64)
