"""CT-mode codec: ct_preprocess / ct_postprocess / latent_CrossEntropy_loss
(models/ct_mcq_vae.py:472-496, 306-311).

CPU part: the oracle restatement against golden vectors minted from the live reference (tests/golden/make_golden_ct.py).
GPU part: the CUDA kernels (through the C ABI) against the goldens and the oracle: indices and one-hots bit-exact, the
loss and its gradient within 1e-5 relative (fp32, north_star).
"""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ct_codec_cfg3", "ct_codec_mcq", "ct_codec_odd"]
TOL = 1e-5


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    g["shape"] = [int(g["B"]), int(g["C"]) * 32, int(g["H"]), int(g["W"])]
    return g


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import ctvq_oracle as O
    g = _load(name)
    N, C = int(g["N"]), int(g["C"])
    assert torch.equal(O.ct_preprocess(g["inds"], g["shape"], N, C), g["onehot"])
    assert torch.equal(O.ct_postprocess(g["scores"], g["shape"], N, C), g["post"])
    assert torch.equal(O.ct_postprocess(g["onehot"], g["shape"], N, C), g["inds"])
    lat = g["latent"].clone().requires_grad_(True)
    loss = O.latent_cross_entropy_loss(lat, g["latent_y"])
    (float(g["g_ce"]) * loss).backward()
    assert torch.equal(loss.detach(), g["ce"])
    assert torch.equal(lat.grad, g["g_latent"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernels_match_golden(name):
    from ct_vae_b200 import ct_codec
    g = _load(name)
    dev = torch.device("cuda:0")
    N, C = int(g["N"]), int(g["C"])
    onehot = ct_codec.ct_preprocess(g["inds"].to(dev), g["shape"], N, C)
    assert onehot.shape == g["onehot"].shape and onehot.dtype == torch.float32 and onehot.is_contiguous()
    assert torch.equal(onehot.cpu(), g["onehot"])
    post = ct_codec.ct_postprocess(g["scores"].to(dev), g["shape"], N, C)
    assert post.dtype == torch.int64 and torch.equal(post.cpu(), g["post"])
    assert torch.equal(ct_codec.ct_postprocess(onehot, g["shape"], N, C).cpu(), g["inds"])   # round trip
    lat = g["latent"].to(dev).requires_grad_(True)
    loss = ct_codec.latent_cross_entropy_loss(lat, g["latent_y"].to(dev))
    (float(g["g_ce"]) * loss).backward()
    assert rel_err(loss.detach().cpu(), g["ce"]) < TOL
    assert rel_err(lat.grad.cpu(), g["g_latent"]) < TOL
    assert torch.equal(lat.grad.cpu() == 0, g["g_latent"] == 0), "clamp mask (latent >= 1e-4) must match"


@pytest.mark.gpu
def test_kernels_at_scale_and_edge_semantics():
    """Bandwidth-sized inputs against the oracle on a sample, a non-contiguous (permuted-view) input as the reference
    produces it, NaN scores (torch.argmax: the first NaN wins) and the out-of-range flag."""
    from ct_vae_b200 import ct_codec
    from oracle import ctvq_oracle as O
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    B, C, H, W, N = 512, 4, 8, 8, 64
    shape = [B, 128, H, W]
    inds = torch.randint(0, N, (B, C, H, W), device=dev)
    onehot = ct_codec.ct_preprocess(inds, shape, N, C)
    assert float(onehot.sum()) == B * C * H * W and torch.equal(onehot.argmax(1).reshape(B, C, H, W), inds)
    assert torch.equal(ct_codec.ct_postprocess(onehot, shape, N, C), inds)
    assert torch.equal(onehot[:4].cpu(), O.ct_preprocess(inds[:4].cpu(), [4, 128, H, W], N, C))
    view = O.ct_preprocess(inds, shape, N, C)            # the reference's non-contiguous permuted view, on the GPU
    assert not view.is_contiguous() and torch.equal(ct_codec.ct_postprocess(view, shape, N, C), inds)
    scores = torch.randn(B, N, C * H, W, device=dev)
    scores[1, 5, 2, 3] = float("nan")
    scores[1, 9, 2, 3] = float("nan")
    scores[2, :, 0, 0] = 1.0
    post = ct_codec.ct_postprocess(scores, shape, N, C)
    assert torch.equal(post.cpu(), O.ct_postprocess(scores.cpu(), shape, N, C))
    lat = (torch.rand(B, N, C * H, W, device=dev) * 0.1).requires_grad_(True)
    lat_y = torch.rand(B, N, C * H, W, device=dev)
    loss = ct_codec.latent_cross_entropy_loss(lat, lat_y)
    loss.backward()
    ref_lat = lat.detach().cpu().requires_grad_(True)
    ref = O.latent_cross_entropy_loss(ref_lat, lat_y.cpu())
    ref.backward()
    assert rel_err(loss.detach().cpu(), ref.detach()) < TOL and rel_err(lat.grad.cpu(), ref_lat.grad) < TOL
    with pytest.raises(RuntimeError):
        ct_codec.ct_preprocess(inds.cpu(), shape, N, C)   # no CPU fallback


def test_ct_codec_install_rebinds_reference_methods():
    import types
    from ct_vae_b200 import ct_codec
    a, b = types.SimpleNamespace(), types.SimpleNamespace()
    ct_codec.install(a, b)
    assert callable(a.ct_preprocess) and callable(a.ct_postprocess) and callable(b.latent_CrossEntropy_loss)


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference_golden(name):
    """The plain-C restatement of the codec (oracle/ctvq_oracle_c.c) against the live-reference goldens: converters
    bit-exact, loss / gradient to 1e-6 (the C oracle sums in double, the reference in fp32)."""
    from oracle import c_oracle as CO
    g = _load(name)
    N, C = int(g["N"]), int(g["C"])
    assert torch.equal(CO.ct_onehot(g["inds"], N), g["onehot"])
    assert torch.equal(CO.ct_class_argmax(g["scores"], C), g["post"])
    assert torch.equal(CO.ct_class_argmax(g["onehot"], C), g["inds"])
    loss, gx = CO.ct_latent_ce(g["latent"], g["latent_y"], float(g["g_ce"]))
    assert rel_err(loss, g["ce"]) < 1e-6
    assert rel_err(gx, g["g_latent"]) < 1e-6
    assert torch.equal(gx == 0, g["g_latent"] == 0)
