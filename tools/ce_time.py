import sys, torch
sys.path.insert(0, ".")
from ct_vae_b200 import ct_codec
from tools.sweep import time_ms
dev = torch.device("cuda:0")
B, C, H, W, N = 16384, 4, 8, 8, 64
lat = (torch.rand(B, N, C * H, W, device=dev) * 0.1).requires_grad_(True)
lat_y = torch.rand(B, N, C * H, W, device=dev)
one = torch.ones((), device=dev)
def ce_fb():
    loss = ct_codec.latent_cross_entropy_loss(lat, lat_y)
    torch.autograd.backward([loss], [one]); lat.grad = None
def ce_f():
    with torch.no_grad():
        return ct_codec.latent_cross_entropy_loss(lat, lat_y)
rows = B * C * H * W
tf, tfb = time_ms(ce_f, 20, None), time_ms(ce_fb, 20, None)
print("CE fwd ms", tf, "GB/s", rows * (8 * N + 12) / tf / 1e6, "fwd+bwd ms", tfb, "GB/s", rows * (16 * N + 24) / tfb / 1e6)
