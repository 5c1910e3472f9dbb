"""CPU-side checks (no GPU compute): the C-ABI library loads and exports every symbol include/ctvq.h
declares; the drop-in modules keep the reference's constructor surface, attribute names, state-dict keys and
seeded initialisation; errors are loud; the injection helper rebinds the reference's names."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import ref_live

HEADER = os.path.join(ROOT, "include", "ctvq.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctvq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ct_vae_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/ctvq.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "python binding table out of sync with the header"


def test_version_and_error_strings():
    from ct_vae_b200 import _lib
    L = _lib.lib()
    assert L.ctvq_version() == 200
    assert b"unsupported" in L.ctvq_strerror(-2)
    assert b"bad argument" in L.ctvq_strerror(-1)
    assert L.ctvq_workspace_bytes(4, 64, 32) >= 64 * 8


def test_bad_arguments_are_rejected_without_a_gpu():
    from ct_vae_b200 import _lib
    L = _lib.lib()
    # null pointers / bad sizes are refused before any CUDA call
    assert L.ctvq_forward(None, None, 1, 4, 4, 1, 4, 8, 1, 0, 0.25, None, None, None, None, None, 0, 0, None) == -1
    assert L.ctvq_reparam_kld_fwd(None, None, None, 1, 1, None, None, None, 0, 0, None) == -1


def test_module_surface_matches_reference_contract():
    import ct_vae_b200 as pkg
    torch.manual_seed(1265)
    vq = pkg.VectorQuantizer(512, 64)
    assert (vq.K, vq.D, vq.beta) == (512, 64, 0.25)
    assert list(vq.state_dict()) == ["embedding.weight"]
    assert float(vq.embedding.weight.abs().max()) <= 1 / 512
    mcq = pkg.MultipleCodebookVectorQuantizer(64, 128, 4, beta=0.25)
    assert mcq.nb_codebooks == 4 and mcq.reduced_embedding_dim == 32
    assert list(mcq.state_dict()) == [f"quantizers.{i}.embedding.weight" for i in range(4)]
    with pytest.raises(AssertionError):
        pkg.MultipleCodebookVectorQuantizer(64, 130, 4)  # models/mcq_vae.py:89


def test_cpu_tensors_raise_loudly():
    import ct_vae_b200 as pkg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.VectorQuantizer(8, 4)(torch.randn(2, 4, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.MultipleCodebookVectorQuantizer(8, 8, 2).compute_inds(torch.randn(2, 8, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.gaussian.reparam_kld(torch.randn(2, 3), torch.randn(2, 3), torch.randn(2, 3))


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "ct_vae_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # no import / include / dlopen / path reference of anything under oracle/ (comments may name it)
                for needle in ("import oracle", "from oracle", "oracle/", "oracle.", "ctvq_oracle", "c_oracle", "ref_live"):
                    assert needle not in text, f"{f} references the oracle ({needle!r})"


@pytest.mark.skipif(not ref_live.available(), reason="reference tree not mounted")
def test_seeded_init_and_keys_equal_the_reference():
    models = ref_live.load()
    import ct_vae_b200 as pkg
    from models.mcq_vae import MultipleCodebookVectorQuantizer as RefMCQ
    from models.vq_vae import VectorQuantizer as RefVQ
    torch.manual_seed(1320)
    a = RefMCQ(64, 128, 4, 0.25)
    torch.manual_seed(1320)
    b = pkg.MultipleCodebookVectorQuantizer(64, 128, 4, 0.25)
    assert list(a.state_dict()) == list(b.state_dict())
    for (_, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y)
    torch.manual_seed(1265)
    a = RefVQ(512, 64)
    torch.manual_seed(1265)
    b = pkg.VectorQuantizer(512, 64)
    assert torch.equal(a.embedding.weight, b.embedding.weight)
    b.load_state_dict(a.state_dict(), strict=True)


@pytest.mark.skipif(not ref_live.available(), reason="reference tree not mounted")
def test_patch_install_and_swap_on_the_reference_models():
    models = ref_live.load()
    import ct_vae_b200 as pkg
    from ct_vae_b200 import patch
    import models.mcq_vae as ref_mcq
    import models.vq_vae as ref_vq
    saved = {(m, n): getattr(m, n) for m in (models, ref_mcq, ref_vq) for n in patch._NAMES if hasattr(m, n)}
    try:
        torch.manual_seed(1320)
        ref_model = models.vae_models["MCQVAE"](in_channels=3, embedding_dim=128, num_embeddings=64,
                                                  hidden_dims=[64, 128, 256], img_size=64, codebooks=4, beta=0.25)
        assert patch.install(models) >= 3
        torch.manual_seed(1320)
        new_model = models.vae_models["MCQVAE"](in_channels=3, embedding_dim=128, num_embeddings=64,
                                                  hidden_dims=[64, 128, 256], img_size=64, codebooks=4, beta=0.25)
        assert isinstance(new_model.vq_layer, pkg.MultipleCodebookVectorQuantizer)
        # identical RNG consumption: every parameter of the patched model equals the reference's
        for (ka, va), (kb, vb) in zip(ref_model.state_dict().items(), new_model.state_dict().items()):
            assert ka == kb and torch.equal(va, vb)
        new_model.load_state_dict(ref_model.state_dict(), strict=True)
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)
    # swap on an existing model keeps the very same Parameter objects
    torch.manual_seed(1265)
    vq_model = models.vae_models["VQVAE"](3, 64, 512)
    w = vq_model.vq_layer.embedding.weight
    new = patch.swap_vq_layer(vq_model)
    assert isinstance(new, pkg.VectorQuantizer) and new.embedding.weight is w
    assert "vq_layer.embedding.weight" in vq_model.state_dict()
