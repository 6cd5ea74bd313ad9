"""Counterpart of vq_ae/layers/vq.py: same classes, constructor signatures, buffers and return
tuples; eval-mode ``forward`` runs the fused sm_100a quantiser kernel.

Training mode (scope row f-4): ``forward`` runs ``_init_ema`` / ``_update_ema`` (vq.py:47-94) as
CUDA kernels (csrc/ema.cu) around the same nearest-code search, with the reference's all-reduce of the
batch statistics under ``torch.distributed``; no autograd graph is built.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from .. import engine as E
from .. import plan as P


class EMAVectorQuantizer(nn.Module):
    """EMA-updated vector quantiser (vq_ae/layers/vq.py:6-154), inference path."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float,
                 decay: float, laplace_alpha: float):
        super().__init__()
        embed = torch.randn(num_embeddings, embedding_dim)
        self.register_buffer("embed", embed)                               # vq.py:27-28
        self.register_buffer("embed_avg", embed.clone())                   # vq.py:29
        self.register_buffer("cluster_size", torch.zeros(num_embeddings))  # vq.py:30
        self.register_buffer("first_pass", torch.as_tensor(1))             # vq.py:34
        self.commitment_cost = commitment_cost
        self.decay = decay
        self.laplace_alpha = laplace_alpha
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        #: number of near-tie vectors (top-2 relative L4 gap < engine.NEAR_TIE_REL_GAP) seen by
        #: the last forward; a 0-dim int32 CUDA tensor (no host sync is forced)
        self.last_near_ties: Optional[torch.Tensor] = None

    def packed(self) -> E.PackedQuantizer:
        return P.packed_quantizer(self)

    # -- reference API -----------------------------------------------------------------------
    def embed_code(self, embed_idx: torch.Tensor) -> torch.Tensor:
        """F.embedding(embed_idx, self.embed) (vq.py:44-45): [...] int -> [..., D]."""
        E.require_cuda(embed_idx, "EMAVectorQuantizer.embed_code")
        bare = E.PackedQuantizer(self.embed, self.commitment_cost)  # table == embed
        flat = embed_idx.reshape(-1)
        out = E.embed_codes(bare, flat, True, 1, flat.numel())
        return out.view(*embed_idx.shape, self.embedding_dim)

    def forward(self, inputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        if self.training:
            # first-pass initialisation + nearest codes + EMA update of the buffers (vq.py:47-94,
            # 118-133); tensors come back detached (no autograd graph is built)
            quantized, idx, loss = P.quantizer_forward_training(self, inputs)
            self.last_near_ties = P.state(self).last_near_ties
            return quantized, idx, loss
        quantized, idx, loss, _ = P.quantizer_forward(self, inputs)
        self.last_near_ties = P.state(self).last_near_ties
        return quantized, idx, loss


class ProjectedEMAVectorQuantizer2d(EMAVectorQuantizer):
    """proj_in (1x1) -> L4 quantiser -> proj_out (1x1) (vq_ae/layers/vq.py:157-192)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float,
                 decay: float, laplace_alpha: float, projection_dim: int):
        super().__init__(num_embeddings, projection_dim, commitment_cost, decay, laplace_alpha)
        self.proj_in, self.proj_out = (
            nn.Conv2d(in_channels=embedding_dim, out_channels=projection_dim, kernel_size=1),
            nn.Conv2d(in_channels=projection_dim, out_channels=embedding_dim, kernel_size=1),
        )

    def decode_codes(self, embed_idx: torch.Tensor, channels_last: bool = False) -> torch.Tensor:
        """proj_out(embed_code(idx)) for stored code maps: [B,H,W] int -> [B,C,H,W]."""
        return P.decode_codes(self, embed_idx, channels_last)
