"""pytest configuration: registers the ``gpu`` marker and shared golden-vector helpers."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names(prefixes=None):
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if prefixes:
        names = [n for n in names if n.startswith(tuple(prefixes))]
    return names


class Golden:
    """One golden case minted from the live reference (tests/golden/make_golden.py)."""

    def __init__(self, name):
        self.name = name
        raw = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.t = {k: torch.from_numpy(np.ascontiguousarray(raw[k])).reshape(raw[k].shape) for k in raw.files}

    def __getitem__(self, k):
        return self.t[k]

    def __contains__(self, k):
        return k in self.t

    @property
    def is_mcq(self):
        return "C" in self.t

    @property
    def codebooks(self):
        if self.is_mcq:
            return [self.t[f"codebook{i}"] for i in range(int(self.t["C"]))]
        return [self.t["codebook"]]

    @property
    def grad_codebooks(self):
        if self.is_mcq:
            return [self.t[f"gE{i}"] for i in range(int(self.t["C"]))]
        return [self.t["gE"]] if "gE" in self.t else None

    @property
    def beta(self):
        return float(self.t["beta"])


QUANT_GOLDENS = [n for n in golden_names() if n not in ("reparam_kld", "vanilla_loss", "gaussian_losses")
                 and not n.startswith(("ct_codec_", "nonfinite_"))]
NONFINITE_GOLDENS = golden_names(["nonfinite_"])  # indices only: rows with NaN / inf / overflowing distances


def rel_err(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
