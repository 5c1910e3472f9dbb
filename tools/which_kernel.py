import sys, torch
sys.path.insert(0, ".")
import ct_vae_b200 as pkg
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
for (K, D, HW, B) in [(512, 64, 256, 4096), (256, 64, 256, 4096), (64, 128, 64, 16384)]:
    m = pkg.VectorQuantizerMS(K, D).to(dev)
    side = int(HW ** 0.5)
    z = torch.randn(B, D, side, side, device=dev, requires_grad=True)
    g = torch.randn(B, D, side, side, device=dev)
    one = torch.ones((), device=dev)
    for _ in range(2):
        o, l = m(z); torch.autograd.backward([o, l], [g, one])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        o, l = m(z); torch.autograd.backward([o, l], [g, one])
        torch.cuda.synchronize()
    print(K, D, [(e.key[:60], round(e.device_time_total)) for e in prof.key_averages() if e.device_time_total > 5])
