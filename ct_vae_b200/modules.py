"""Drop-in nn.Modules with the reference's exact constructor / forward / compute_inds / compute_latents
signatures, attribute names and state-dict keys, running on the ctvq CUDA kernels.

    reference class                                   this module
    models/vq_vae.py:7-55    VectorQuantizer          VectorQuantizer
    models/mcq_vae.py:7-75   VectorQuantizerMS        VectorQuantizerMS
    models/mcq_vae.py:78-137 MultipleCodebookVectorQuantizer   MultipleCodebookVectorQuantizer

Parameters are ordinary ``nn.Embedding`` weights created and initialised by the same calls in the same
order as the reference (``nn.Embedding(K, D)`` then ``uniform_(-1/K, 1/K)``), so a seeded construction
yields identical initial codebooks and ``state_dict()`` keys (``embedding.weight``,
``quantizers.{i}.embedding.weight``) load either way (run.py:86-89).
"""
from typing import Optional

import torch
from torch import nn

from . import functional as F_

Tensor = torch.Tensor


class _NearTieCounter:
    """Near-tie accounting (BASELINE.json north_star: near-ties with a relative top-2 distance gap below 1e-6 are
    "counted and reported, not hidden").  Every argmin launch ADDS its near-tie rows to one int64 device counter owned by
    the module -- a plain attribute, not a buffer: state_dict keys stay the reference's, DDP does not broadcast it, and
    nothing synchronises until ``near_tie_rows()`` is read."""

    _near_tie: Optional[Tensor] = None
    count_near_ties = True

    def _counter(self, latents: Tensor) -> Optional[Tensor]:
        if not self.count_near_ties or not latents.is_cuda:
            return None
        c = self._near_tie
        if c is None or c.device != latents.device:
            c = F_.new_near_tie_counter(latents.device)
            self._near_tie = c
        return c

    def near_tie_rows(self, reset: bool = False) -> int:
        """(row, codebook) pairs quantised so far whose top-2 distance gap was below 1e-6 relative (synchronises)."""
        c = self._near_tie
        n = int(c.item()) if c is not None else 0
        if reset and c is not None:
            c.zero_()
        return n


class VectorQuantizer(_NearTieCounter, nn.Module):
    """models/vq_vae.py:7-55."""

    def __init__(self, num_embeddings: int, embedding_dim: int, beta: float = 0.25):
        super().__init__()
        self.K = num_embeddings
        self.D = embedding_dim
        self.beta = beta
        self.embedding = nn.Embedding(self.K, self.D)
        self.embedding.weight.data.uniform_(-1 / self.K, 1 / self.K)
        self.grad_comm = None  # optional ct_vae_b200.dist.CodebookGradComm

    def forward(self, latents: Tensor):
        out, loss, _, _ = F_.quantize(latents, [self.embedding.weight], self.beta, comm=self.grad_comm,
                                      counter=self._counter(latents))
        return out, loss  # [B x D x H x W], 0-d vq_loss


class VectorQuantizerMS(VectorQuantizer):
    """models/mcq_vae.py:7-75: index computation separated from the quantisation step."""

    def compute_inds(self, latents: Tensor) -> Tensor:
        (inds,) = F_.compute_inds([latents], [self.embedding.weight], counter=self._counter(latents))
        return inds[:, 0]  # [B x H x W] int64

    def compute_latents(self, latents: Tensor, encoding_inds: Tensor):
        out, loss, _, _ = F_.quantize(latents, [self.embedding.weight], self.beta, inds=encoding_inds,
                                      comm=self.grad_comm)
        return out, loss

    def forward(self, latents: Tensor, inds: bool = False):
        out, loss, encoding_inds, _ = F_.quantize(latents, [self.embedding.weight], self.beta, comm=self.grad_comm,
                                                  counter=self._counter(latents))
        if inds:
            return out, loss, encoding_inds[:, 0]
        return out, loss


class MultipleCodebookVectorQuantizer(_NearTieCounter, nn.Module):
    """models/mcq_vae.py:78-137: C codebooks sharing the embedding dimension, ONE launch for all of them.

    ``chan_stride`` is 1 to reproduce the reference's ``latents[:, i:i+d]`` slicing (models/mcq_vae.py:104,117:
    codebook i reads channels i .. i+d-1, so slices overlap); set the attribute to ``reduced_embedding_dim``
    for disjoint slices.
    """

    chan_stride = 1

    def __init__(self, num_embeddings: int, embedding_dim: int, codebooks: int, beta: float = 0.25):
        super().__init__()
        assert embedding_dim % codebooks == 0  # embedding size must be divided between all codebooks
        self.nb_codebooks = codebooks
        self.reduced_embedding_dim = embedding_dim // codebooks
        self.quantizers = nn.ModuleList([VectorQuantizerMS(num_embeddings, self.reduced_embedding_dim, beta)
                                         for _ in range(codebooks)])
        self.grad_comm = None

    def _weights(self):
        return [q.embedding.weight for q in self.quantizers]

    @property
    def beta(self) -> float:
        return self.quantizers[0].beta

    def compute_inds(self, latents: Tensor) -> Tensor:
        (inds,) = F_.compute_inds([latents], self._weights(), self.chan_stride, counter=self._counter(latents))
        return inds  # [B x C x H x W]

    def compute_inds_pair(self, latents_x: Tensor, latents_y: Tensor):
        """Both members of a transition pair in one launch (models/ct_mcq_vae.py:530,536,555-556)."""
        return tuple(F_.compute_inds([latents_x, latents_y], self._weights(), self.chan_stride,
                                     counter=self._counter(latents_x)))

    def compute_latents(self, latents: Tensor, encoding_inds: Tensor):
        out, loss, _, _ = F_.quantize(latents, self._weights(), self.beta, self.chan_stride, inds=encoding_inds,
                                      comm=self.grad_comm)
        return out, loss

    def forward(self, latents: Tensor, inds: bool = False):
        out, loss, encoding_inds, _ = F_.quantize(latents, self._weights(), self.beta, self.chan_stride,
                                                  comm=self.grad_comm, counter=self._counter(latents))
        if inds:
            return out, loss, encoding_inds
        return out, loss


def near_tie_rows(module: nn.Module, reset: bool = False) -> int:
    """Sum of the near-tie counters of every quantiser inside ``module`` (one host sync per quantiser)."""
    return sum(m.near_tie_rows(reset) for m in module.modules() if isinstance(m, _NearTieCounter))


def attach_grad_comm(module: nn.Module, comm: Optional[object]) -> int:
    """Give every quantiser inside ``module`` the communicator its backward all-reduces codebook grads on."""
    n = 0
    for m in module.modules():
        if isinstance(m, (VectorQuantizer, MultipleCodebookVectorQuantizer)):
            m.grad_comm = comm
            n += 1
    return n
