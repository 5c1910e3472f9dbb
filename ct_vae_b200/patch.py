"""Injection of the drop-in quantisers into an UNMODIFIED checkout of the reference.

    import models                      # the reference package
    import ct_vae_b200.patch as patch
    patch.install(models)              # before vae_models[...](**cfg) is called (run.py:52)

or, for an already-built model::

    patch.swap_vq_layer(model)         # replaces model.vq_layer, sharing its Parameters
"""
import sys
from typing import Optional

from torch import nn

from . import modules

_NAMES = ("VectorQuantizer", "VectorQuantizerMS", "MultipleCodebookVectorQuantizer")


def install(models_pkg=None) -> int:
    """Rebind the three quantiser class names in the reference's ``models`` package and its sub-modules
    (models/__init__.py:21,24 star-imports them; vq_vae.py:124, mcq_vae.py:196, ct_mcq_vae.py:400 look
    them up by module-global name at construction time).  Returns the number of bindings replaced."""
    if models_pkg is None:
        models_pkg = sys.modules.get("models")
        if models_pkg is None:
            raise RuntimeError("import the reference's `models` package before patch.install()")
    n = 0
    targets = [models_pkg] + [m for name, m in list(sys.modules.items())
                              if m is not None and name.startswith(models_pkg.__name__ + ".")]
    for mod in targets:
        for name in _NAMES:
            if hasattr(mod, name):
                setattr(mod, name, getattr(modules, name))
                n += 1
    return n


def swap_vq_layer(model: nn.Module, attr: str = "vq_layer") -> nn.Module:
    """Replace ``model.<attr>`` (a reference quantiser) by the drop-in, re-using the SAME Parameter objects so
    optimiser state, DDP registration and checkpoints are unaffected."""
    old = getattr(model, attr)
    kind = type(old).__name__
    if kind == "MultipleCodebookVectorQuantizer":
        k, d = old.quantizers[0].embedding.weight.shape
        new = modules.MultipleCodebookVectorQuantizer(k, d * old.nb_codebooks, old.nb_codebooks, old.quantizers[0].beta)
        for nq, oq in zip(new.quantizers, old.quantizers):
            nq.embedding.weight = oq.embedding.weight
    elif kind in ("VectorQuantizer", "VectorQuantizerMS"):
        k, d = old.embedding.weight.shape
        new = getattr(modules, kind)(k, d, old.beta)
        new.embedding.weight = old.embedding.weight
    else:
        raise TypeError(f"{attr} is a {kind}, not a reference quantiser")
    new.train(old.training)
    setattr(model, attr, new)
    return new
