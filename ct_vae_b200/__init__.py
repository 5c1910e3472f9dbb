"""ct_vae_b200 — B200-native (sm_100a) codebook quantiser for CT-VAE / MCQ-VAE / VQ-VAE.

Host side: Python/PyTorch (device memory, streams, autograd, torch.distributed plumbing).
Product: libctvq.so — hand-written CUDA kernels behind the C ABI in include/ctvq.h.
"""
from .modules import (MultipleCodebookVectorQuantizer, VectorQuantizer, VectorQuantizerMS,  # noqa: F401
                      attach_grad_comm, near_tie_rows)
from ._lib import raise_if_bad_indices, set_validate  # noqa: F401
from . import ct_codec, functional, gaussian, patch  # noqa: F401

__version__ = "0.2.0"
