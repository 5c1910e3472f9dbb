"""Injection of the drop-in quantisers into an UNMODIFIED checkout of the reference.

    import models                      # the reference package
    import ct_vae_b200.patch as patch
    patch.install(models)              # before vae_models[...](**cfg) is called (run.py:52)

or, for an already-built model::

    patch.swap_vq_layer(model)         # replaces model.vq_layer, sharing its Parameters
"""
import contextlib
import functools
import inspect
import sys
from typing import Optional

from torch import nn

from . import modules

_NAMES = ("VectorQuantizer", "VectorQuantizerMS", "MultipleCodebookVectorQuantizer")
_PAIR_METHODS = ("forward_action", "forward_causal")


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


# ----------------------------------------------------------------------------------------------------------------------
# transition-pair batching at the reference's own call sites
# ----------------------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def _memoised(model, encodings, indices):
    """While active, ``model.encode(t)`` / ``model.vq_layer.compute_inds(t)`` answer from the given
    ``{id(tensor): (tensor, result)}`` tables (identity-checked) and fall through to the real methods otherwise."""
    vq = model.vq_layer
    real_encode, real_inds = model.encode, vq.compute_inds

    def encode(t, *a, **k):
        hit = encodings.get(id(t))
        return hit[1] if hit is not None and hit[0] is t and not a and not k else real_encode(t, *a, **k)

    def compute_inds(t, *a, **k):
        hit = indices.get(id(t))
        return hit[1] if hit is not None and hit[0] is t and not a and not k else real_inds(t, *a, **k)

    model.encode, vq.compute_inds = encode, compute_inds  # instance attributes shadow the class methods
    try:
        yield
    finally:
        del model.encode
        del vq.compute_inds


def _pair_batched(orig):
    """Wrap ``CTMCQVAE.forward_action`` / ``forward_causal`` (models/ct_mcq_vae.py:525-567).  Both quantise the encoder
    latents of ``input`` AND of ``input_y`` with two separate ``vq_layer.compute_inds`` calls (:530 + :536, :555-556);
    the wrapper encodes both images first (same order as the reference: x, then y, so BatchNorm statistics see the same
    sequence), quantises the pair in ONE launch (``compute_inds_pair``, n_seg = 2 in ctvq_argmin) and then runs the
    reference's OWN method body unchanged, with ``encode`` / ``compute_inds`` answering from those results."""
    sig = inspect.signature(orig)

    @functools.wraps(orig)
    def wrapped(self, *args, **kwargs):
        vq = getattr(self, "vq_layer", None)
        try:
            bound = sig.bind(self, *args, **kwargs).arguments
        except TypeError:
            return orig(self, *args, **kwargs)
        x, y = bound.get("input"), bound.get("input_y")
        if x is None or y is None or not hasattr(vq, "compute_inds_pair"):
            return orig(self, *args, **kwargs)  # nothing to pair (the reference fails by itself on input_y=None)
        y = y.to(x.device)
        enc_x = self.encode(x)
        enc_y = self.encode(y)
        ix, iy = vq.compute_inds_pair(enc_x[0], enc_y[0])
        # the reference's body calls encode(input_y) with the ORIGINAL object: register both it and the moved copy
        y_orig = bound.get("input_y")
        enc_tab = {id(x): (x, enc_x), id(y): (y, enc_y), id(y_orig): (y_orig, enc_y)}
        ind_tab = {id(enc_x[0]): (enc_x[0], ix), id(enc_y[0]): (enc_y[0], iy)}
        with _memoised(self, enc_tab, ind_tab):
            return orig(self, *args, **kwargs)

    wrapped._ctvq_pair_batched = True
    wrapped.__wrapped__ = orig
    return wrapped


def pair_batch_class(cls) -> int:
    """Rebind ``forward_action`` / ``forward_causal`` of a CTMCQVAE-shaped class (and the entries of its
    ``FORWARD_MODES`` dispatch table, models/ct_mcq_vae.py:570-574, which holds the function objects themselves).
    Idempotent.  Returns the number of methods wrapped."""
    n = 0
    for name in _PAIR_METHODS:
        orig = cls.__dict__.get(name)
        if orig is None or getattr(orig, "_ctvq_pair_batched", False):
            continue
        new = _pair_batched(orig)
        setattr(cls, name, new)
        table = cls.__dict__.get("FORWARD_MODES")
        if isinstance(table, dict):
            for mode, fn in list(table.items()):
                if fn is orig:
                    table[mode] = new
        n += 1
    return n


def install_ct(models_pkg=None) -> int:
    """``install()`` plus the pair batching of ``CTMCQVAE`` in an unmodified checkout: after this call
    ``CTMCQVAE.forward(..., mode="action" | "causal")`` issues ONE argmin launch for the (x, y) pair instead of two
    (models/ct_mcq_vae.py:530,536,555-556).  Returns the number of bindings replaced."""
    if models_pkg is None:
        models_pkg = sys.modules.get("models")
        if models_pkg is None:
            raise RuntimeError("import the reference's `models` package before patch.install_ct()")
    n = install(models_pkg)
    seen = set()
    for mod in [models_pkg] + [m for name, m in list(sys.modules.items())
                               if m is not None and name.startswith(models_pkg.__name__ + ".")]:
        cls = getattr(mod, "CTMCQVAE", None)
        if inspect.isclass(cls) and id(cls) not in seen:
            seen.add(id(cls))
            n += pair_batch_class(cls)
    return n
