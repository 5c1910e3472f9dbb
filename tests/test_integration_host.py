"""Host-side (CPU) tests of the injection helpers against the REAL reference classes (skipped where /root/reference is
not mounted): pair batching of CTMCQVAE.forward_action / forward_causal (models/ct_mcq_vae.py:525-567)
and the fused-Gaussian rebinding of VanillaVAE / BetaVAE (models/vanilla_vae.py:107-146, models/beta_vae.py:112-152,
SURVEY §8 a13-a14).

The product ops refuse CPU tensors (no fallback), so these tests stand the kernels in with the ORACLE (allowed: tests/
is one of the places that may use it) to check the HOST logic: which entry point is reached how often, argument
plumbing, dict keys, the capacity schedule.  The kernels themselves are checked in the -m gpu tests."""
import pytest
import torch
from torch import nn

from oracle import ctvq_oracle as O
from oracle import ref_live

pytestmark = pytest.mark.skipif(not ref_live.available(), reason="reference tree not mounted")


class _CountingQuantiser(nn.Module):
    """CPU stand-in with the drop-in's surface (compute_inds, compute_inds_pair, compute_latents), oracle arithmetic."""

    def __init__(self, K=16, D=8):
        super().__init__()
        torch.manual_seed(3)
        self.book = nn.Parameter(torch.randn(K, D) * 0.5)
        self.calls = {"compute_inds": 0, "compute_inds_pair": 0, "compute_latents": 0}

    def compute_inds(self, latents):
        self.calls["compute_inds"] += 1
        return O.mcq_compute_inds(latents.detach(), [self.book.detach()])

    def compute_inds_pair(self, x, y):
        self.calls["compute_inds_pair"] += 1
        return (O.mcq_compute_inds(x.detach(), [self.book.detach()]), O.mcq_compute_inds(y.detach(), [self.book.detach()]))

    def compute_latents(self, latents, inds):
        self.calls["compute_latents"] += 1
        out, loss, _ = O.mcq_compute_latents(latents, inds, [self.book], 0.25)
        return out, loss


class _FakeTransition:
    """The calls CTMCQVAE makes on its ct_layer (the real CausalTransition needs torch_geometric)."""

    def __call__(self, onehot):
        return onehot * 0.5 + 0.01, torch.tensor(0.125), {"m": torch.tensor(0.0)}

    def forward_action(self, onehot, action):
        return onehot * 0.5 + 0.01, torch.tensor(0.125), {"m": torch.tensor(1.0)}

    def forward_transition(self, ox, oy):
        return (ox - oy).abs().mean(dim=(1, 2, 3)).unsqueeze(-1).repeat(1, 3), torch.tensor(0.25), {"m": torch.tensor(2.0)}

    def latent_loss(self, a, b):
        return ((a - b) ** 2).mean()

    def causal_accuracy(self, recons_action, action):
        return torch.tensor(0.5)

    def causal_undirected_accuracy(self, recons_action, action):
        return torch.tensor(0.75)


def _stand_in(models):
    cls = models.ct_mcq_vae.CTMCQVAE
    me = cls.__new__(cls)
    nn.Module.__init__(me)
    torch.manual_seed(0)
    me.encoder = nn.Sequential(nn.Conv2d(3, 8, 4, 4), nn.BatchNorm2d(8), nn.LeakyReLU())
    me.decoder = nn.Sequential(nn.ConvTranspose2d(8, 3, 4, 4), nn.Tanh())
    me.vq_layer = _CountingQuantiser(16, 8)
    me.ct_layer = _FakeTransition()
    me.num_embeddings, me.codebooks, me.skip_transition, me.gamma = 16, 1, False, 0.25
    return me


@pytest.fixture()
def models():
    return ref_live.load()


def _flat(res):
    out = []
    for r in res:
        if isinstance(r, dict):
            out += [v for _, v in sorted(r.items()) if isinstance(v, torch.Tensor)]
        elif isinstance(r, torch.Tensor):
            out.append(r)
    return out


@pytest.mark.parametrize("mode", ["action", "causal", "base"])
def test_install_ct_routes_the_pair_through_one_launch(models, mode):
    import ct_vae_b200.patch as patch
    cls = models.ct_mcq_vae.CTMCQVAE
    originals = {n: cls.__dict__[n] for n in ("forward_action", "forward_causal")}
    table = dict(cls.FORWARD_MODES)
    torch.manual_seed(1)
    x, y = torch.rand(4, 3, 16, 16), torch.rand(4, 3, 16, 16)
    action = torch.nn.functional.one_hot(torch.randint(0, 3, (4,)), 3).float()
    plain = _stand_in(models)
    ref_res = plain(x, input_y=y, action=action, mode=mode)       # the unpatched reference method bodies
    ref_calls = dict(plain.vq_layer.calls)
    try:
        assert patch.pair_batch_class(cls) == 2 and patch.pair_batch_class(cls) == 0   # idempotent
        assert cls.FORWARD_MODES["action"] is cls.__dict__["forward_action"], "the dispatch table must be rebound too"
        me = _stand_in(models)
        res = me(x, input_y=y, action=action, mode=mode)
        calls = me.vq_layer.calls
        if mode == "base":
            assert calls == ref_calls and calls["compute_inds_pair"] == 0
        else:
            assert ref_calls["compute_inds"] == 2 and ref_calls["compute_inds_pair"] == 0
            assert calls["compute_inds"] == 0 and calls["compute_inds_pair"] == 1, calls
            assert calls["compute_latents"] == ref_calls["compute_latents"]
        for a, b in zip(_flat(res), _flat(ref_res)):
            assert torch.equal(a, b), "pair batching must not change a single output"
        assert "encode" not in me.__dict__ and "compute_inds" not in me.vq_layer.__dict__, "memo must be removed"
        # BatchNorm saw x then y in both runs: identical running statistics
        assert torch.equal(me.encoder[1].running_mean, plain.encoder[1].running_mean)
        # loss_function consumes the result unchanged (dict keys of models/ct_mcq_vae.py:594-620)
        out = me.loss_function(*res)
        assert {"loss", "Reconstruction_Loss", "VQ_Loss", "CT_Loss"} <= set(out)
    finally:
        for n, fn in originals.items():
            setattr(cls, n, fn)
        cls.FORWARD_MODES.clear()
        cls.FORWARD_MODES.update(table)


def test_install_ct_rebinds_names_and_methods(models):
    import ct_vae_b200 as pkg
    import ct_vae_b200.patch as patch
    cls = models.ct_mcq_vae.CTMCQVAE
    originals = {n: cls.__dict__[n] for n in ("forward_action", "forward_causal")}
    table = dict(cls.FORWARD_MODES)
    saved = {(m, n): getattr(m, n) for m in (models, models.vq_vae, models.mcq_vae, models.ct_mcq_vae)
             for n in patch._NAMES if hasattr(m, n)}
    try:
        n = patch.install_ct(models)
        assert n >= len(saved) + 2
        assert models.mcq_vae.MultipleCodebookVectorQuantizer is pkg.MultipleCodebookVectorQuantizer
        assert getattr(cls.forward_causal, "_ctvq_pair_batched", False)
        # a model built AFTER install gets the drop-in quantiser (models/mcq_vae.py:196 looks the name up at build time)
        m = models.MCQVAE(3, 128, 64, hidden_dims=[16, 32], codebooks=4)
        assert isinstance(m.vq_layer, pkg.MultipleCodebookVectorQuantizer)
    finally:
        for (m, nme), v in saved.items():
            setattr(m, nme, v)
        for nme, fn in originals.items():
            setattr(cls, nme, fn)
        cls.FORWARD_MODES.clear()
        cls.FORWARD_MODES.update(table)


def _oracle_reparam_kld(mu, logvar, eps=None):
    if eps is None:
        eps = torch.randn_like(logvar)
    return O.reparameterize(mu, logvar, eps), O.kld(mu, logvar)


@pytest.mark.parametrize("kind", ["vanilla", "betaH", "betaB"])
def test_gaussian_install_on_the_real_classes(models, kind, monkeypatch):
    """Patched VanillaVAE / BetaVAE forward + loss_function == the unpatched reference, value for value, with the
    kernel stood in by the oracle (CPU): same eps stream, same dict, same num_iter / capacity schedule."""
    from ct_vae_b200 import gaussian
    cls = models.VanillaVAE if kind == "vanilla" else models.BetaVAE
    saved = {n: cls.__dict__[n] for n in ("reparameterize", "loss_function")}
    kw = {} if kind == "vanilla" else dict(beta=4, gamma=10.0, max_capacity=25, Capacity_max_iter=3, loss_type=kind[-1])

    def run(patched):
        models.BetaVAE.num_iter = 0
        torch.manual_seed(5)
        m = cls(3, 16, hidden_dims=[8, 16, 32, 64, 512], **kw)
        x = torch.rand(4, 3, 64, 64)
        outs = []
        for it in range(4):
            torch.manual_seed(100 + it)
            res = m(x)
            d = m.loss_function(*res, M_N=0.005)
            d["loss"].sum().backward()
            outs.append({k: v.detach().clone() for k, v in d.items()})
        grads = [p.grad.clone() for p in m.parameters() if p.grad is not None]
        return outs, grads, getattr(m, "num_iter", None)

    ref_outs, ref_grads, ref_iter = run(False)
    monkeypatch.setattr(gaussian, "reparam_kld", _oracle_reparam_kld)
    try:
        assert gaussian.install(cls) == 1 and gaussian.install(cls) == 0
        outs, grads, n_iter = run(True)
    finally:
        for n, fn in saved.items():
            setattr(cls, n, fn)
        del cls._ctvq_fused_gaussian
    assert n_iter == ref_iter
    for a, b in zip(outs, ref_outs):
        assert set(a) == set(b)
        for k in a:
            assert a[k].shape == b[k].shape
            assert torch.allclose(a[k], b[k], rtol=1e-6, atol=0), (kind, k)
    for a, b in zip(grads, ref_grads):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7)
