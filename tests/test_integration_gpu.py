"""GPU tests of the injection helpers on golden data minted from the live reference.  The reference
tree does not travel to the GPU box, so the model classes here are SHELLS with the reference's call structure
(models/vq_vae.py:189-211, models/ct_mcq_vae.py:525-567, models/vanilla_vae.py:119-146, models/beta_vae.py:124-152);
the real classes are patched and compared value for value on the CPU in tests/test_integration_host.py."""
import types

import pytest
import torch
from torch import nn
from torch.nn import functional as F

from conftest import Golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _fake_models_package():
    """A module object shaped like the reference's `models` package: the three quantiser names (stand-ins that must be
    REPLACED by install) and a VQVAE-like shell that looks `VectorQuantizer` up at construction time."""
    pkg = types.ModuleType("fake_models")

    class _Refuse(nn.Module):
        def __init__(self, *a, **k):
            raise AssertionError("the stock quantiser was constructed: install() did not rebind the name")

    pkg.VectorQuantizer = pkg.VectorQuantizerMS = pkg.MultipleCodebookVectorQuantizer = _Refuse

    class VQVAEShell(nn.Module):
        def __init__(self, embedding_dim, num_embeddings, beta=0.25):
            super().__init__()
            self.vq_layer = pkg.VectorQuantizer(num_embeddings, embedding_dim, beta)  # models/vq_vae.py:124-126
            self.scale = nn.Parameter(torch.ones(()))  # a one-parameter "encoder" so gradients reach something upstream

        def encode(self, x):
            return [x * self.scale]

        def decode(self, z):
            return z

        def forward(self, x):  # models/vq_vae.py:189-192
            encoding = self.encode(x)[0]
            quantized, vq_loss = self.vq_layer(encoding)
            return [self.decode(quantized), x, vq_loss]

        def loss_function(self, *args, **kwargs):  # models/vq_vae.py:194-211
            recons, inp, vq_loss = args[0], args[1], args[2]
            recons_loss = F.mse_loss(recons, inp)
            return {"loss": recons_loss + vq_loss, "Reconstruction_Loss": recons_loss, "VQ_Loss": vq_loss}

    pkg.VQVAEShell = VQVAEShell
    return pkg


def test_patched_shell_forward_loss_backward_on_the_reference_recipe():
    """tests/test_vq_vae.py:17-29 recipe (VQVAE(3,64,512), randn(16,3,64,64)): the golden holds the real encoder's
    latents, the reference's quantised output, vq_loss, Reconstruction_Loss and model loss.  A shell built through
    patch.install() must reproduce the quantiser part exactly and `recons_loss + vq_loss == model_loss`."""
    import ct_vae_b200 as pkg
    import ct_vae_b200.patch as patch
    from oracle import ctvq_oracle as O
    g = Golden("vqvae_recipe_encoder_latents")
    dev = torch.device("cuda:0")
    fake = _fake_models_package()
    assert patch.install(fake) == 3
    model = fake.VQVAEShell(64, 512).to(dev)
    assert isinstance(model.vq_layer, pkg.VectorQuantizer)
    assert list(model.state_dict()) == ["scale", "vq_layer.embedding.weight"]  # run.py:86-89 checkpoint keys
    model.vq_layer.embedding.weight.data.copy_(g["codebook"])
    x = g["z"].to(dev)
    res = model(x)
    out = model.loss_function(*res, M_N=0.005)
    assert rel_err(res[2].detach().cpu(), g["loss"]) < TOL
    assert torch.equal(res[0].detach().cpu(), g["out"])
    assert rel_err((g["recons_loss"] + res[2].detach().cpu()), g["model_loss"]) < TOL
    out["loss"].backward()
    torch.cuda.synchronize()
    # gradients against the oracle's explicit backward: d(mse(out, x) + vq_loss)
    z_cpu, e_cpu = g["z"], g["codebook"]
    inds = O.vq_compute_inds(z_cpu, e_cpu)
    o_out, _ = O.vq_compute_latents(z_cpu, inds, e_cpu, 0.25)
    g_out = 2.0 * (o_out - z_cpu) / z_cpu.numel()
    gz, ge = O.vq_backward(z_cpu, inds, e_cpu, 0.25, g_out, torch.tensor(1.0))
    assert rel_err(model.vq_layer.embedding.weight.grad.cpu(), ge) < TOL
    exp_scale_grad = (gz.double() * z_cpu.double()).sum()  # d/dscale through encode(x) = x * scale
    assert abs(float(model.scale.grad) - float(exp_scale_grad)) < 1e-3 * max(1e-3, abs(float(exp_scale_grad)))
    assert model.vq_layer.near_tie_rows() >= 0


class _CTShell(nn.Module):
    """Call structure of CTMCQVAE.forward_action / forward_causal (models/ct_mcq_vae.py:525-567) around the quantiser."""

    def __init__(self, vq):
        super().__init__()
        torch.manual_seed(0)
        self.encoder = nn.Sequential(nn.Conv2d(3, 128, 8, 8), nn.LeakyReLU())
        self.vq_layer = vq
        self.num_embeddings, self.codebooks = 64, 1

    def encode(self, x):
        return [self.encoder(x)]

    def forward_action(self, input, action, input_y=None, **kwargs):
        latents = self.encode(input)[0]
        encoding_inds = self.vq_layer.compute_inds(latents)
        inds_y = self.vq_layer.compute_inds(self.encode(input_y)[0])
        quantized, _ = self.vq_layer.compute_latents(latents, encoding_inds)
        return [quantized, encoding_inds, inds_y]

    def forward_causal(self, input, input_y, action=None, **kwargs):
        latents_x = self.encode(input)[0]
        latents_y = self.encode(input_y)[0]
        return [self.vq_layer.compute_inds(latents_x), self.vq_layer.compute_inds(latents_y)]

    FORWARD_MODES = {"action": forward_action, "causal": forward_causal}

    def forward(self, input, input_y=None, action=None, mode="action"):
        return _CTShell.FORWARD_MODES[mode](self, input=input, input_y=input_y, action=action)


@pytest.mark.parametrize("mode", ["action", "causal"])
def test_pair_batching_on_gpu_is_one_launch_and_identical(mode):
    """configs/ct_mcq_vae.yaml quantiser (C=1, d=128, K=64, latents [16,128,8,8]): the patched methods reach
    ctvq_argmin ONCE with n_seg = 2 and return exactly what two separate launches return."""
    import ct_vae_b200 as pkg
    import ct_vae_b200.patch as patch
    from ct_vae_b200 import functional as F_
    dev = torch.device("cuda:0")
    torch.manual_seed(1250)
    vq = pkg.MultipleCodebookVectorQuantizer(64, 128, 1, 0.1)
    vq.quantizers[0].embedding.weight.data = torch.randn(64, 128) * 0.5
    model = _CTShell(vq).to(dev)
    x, y = torch.rand(16, 3, 64, 64, device=dev), torch.rand(16, 3, 64, 64, device=dev)
    ref = [t.clone() for t in model(x, input_y=y, mode=mode)]
    calls = []
    real = F_.compute_inds

    def spy(latents_list, *a, **k):
        calls.append(len(latents_list))
        return real(latents_list, *a, **k)

    saved = {n: _CTShell.__dict__[n] for n in ("forward_action", "forward_causal")}
    table = dict(_CTShell.FORWARD_MODES)
    try:
        assert patch.pair_batch_class(_CTShell) == 2
        F_.compute_inds = spy
        got = model(x, input_y=y, mode=mode)
    finally:
        F_.compute_inds = real
        for n, fn in saved.items():
            setattr(_CTShell, n, fn)
        _CTShell.FORWARD_MODES.clear()
        _CTShell.FORWARD_MODES.update(table)
    assert calls == [2], f"expected ONE argmin launch over the (x, y) pair, saw segments per launch = {calls}"
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


class _VanillaShell(nn.Module):
    def reparameterize(self, mu, logvar):
        raise AssertionError("stock reparameterize reached")

    def loss_function(self, *args, **kwargs):
        raise AssertionError("stock loss_function reached")


class _BetaShell(nn.Module):
    num_iter = 0

    def __init__(self, beta=4, gamma=1000.0, max_capacity=25, Capacity_max_iter=1e5, loss_type="B"):
        super().__init__()
        self.beta, self.gamma, self.loss_type = beta, gamma, loss_type
        self.C_max = torch.Tensor([max_capacity])
        self.C_stop_iter = Capacity_max_iter

    def reparameterize(self, mu, logvar):
        raise AssertionError("stock reparameterize reached")

    def loss_function(self, *args, **kwargs):
        raise AssertionError("stock loss_function reached")


def test_gaussian_install_against_the_reference_goldens():
    """VanillaVAE / BetaVAE-H / BetaVAE-B loss dicts (models/vanilla_vae.py:139-146, models/beta_vae.py:139-152) from the
    fused kernel == the live reference's values, five successive calls (capacity schedule saturating at call 3)."""
    from ct_vae_b200 import gaussian
    from oracle import ctvq_oracle as O
    g = Golden("gaussian_losses")
    dev = torch.device("cuda:0")
    assert gaussian.install(_VanillaShell, _BetaShell) == 2
    mu0, lv0 = g["mu"].to(dev), g["logvar"].to(dev)
    rec, inp, m_n = g["recons"].to(dev), g["input"].to(dev), float(g["M_N"])
    # Vanilla: z from the fused op with the generator at a known position == oracle on the same eps
    v = _VanillaShell()
    mu, lv = mu0.clone().requires_grad_(True), lv0.clone().requires_grad_(True)
    torch.manual_seed(9)
    z = v.reparameterize(mu, lv)
    torch.manual_seed(9)
    eps = torch.randn_like(lv0)
    assert rel_err(z.detach().cpu(), O.reparameterize(g["mu"], g["logvar"], eps.cpu())) < TOL
    d = v.loss_function(rec, inp, mu, lv, M_N=m_n)
    assert set(d) == {"loss", "Reconstruction_Loss", "KLD"}
    assert rel_err(d["loss"].detach().cpu(), g["vanilla_loss"]) < TOL
    assert rel_err(d["KLD"].cpu(), g["vanilla_KLD"]) < TOL and not d["KLD"].requires_grad
    (d["loss"] + (z * 0.01).sum()).backward()  # ONE backward kernel carries both g_z and g_kld
    g_mu, g_lv = O.reparam_kld_backward(g["mu"], g["logvar"], eps.cpu(), torch.full_like(g["mu"], 0.01), torch.tensor(m_n))
    assert rel_err(mu.grad.cpu(), g_mu) < TOL and rel_err(lv.grad.cpu(), g_lv) < TOL
    # loss_function on tensors that never went through reparameterize still gives the KL term
    d2 = v.loss_function(rec, inp, mu0, lv0, M_N=m_n)
    assert rel_err(d2["loss"].cpu(), g["vanilla_loss"]) < TOL
    for lt in ("H", "B"):
        _BetaShell.num_iter = 0
        b = _BetaShell(beta=float(g["beta"]), gamma=float(g["gamma"]), max_capacity=float(g["max_capacity"]),
                       Capacity_max_iter=float(g["Capacity_max_iter"]), loss_type=lt)
        for it in range(1, 6):
            b.reparameterize(mu0, lv0)
            d = b.loss_function(rec, inp, mu0, lv0, M_N=m_n)
            assert b.num_iter == it
            assert d["loss"].shape == g[f"beta{lt}_loss_{it}"].shape
            assert rel_err(d["loss"].cpu(), g[f"beta{lt}_loss_{it}"]) < TOL, (lt, it)
            assert rel_err(d["KLD"].cpu(), g[f"beta{lt}_KLD_{it}"]) < TOL
            assert rel_err(d["Reconstruction_Loss"].cpu(), g[f"beta{lt}_recons_{it}"]) < TOL
