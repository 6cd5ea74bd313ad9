"""Counterpart of vq_ae/layers/vq.py: same classes, constructor signatures, buffers and return
tuples; eval-mode ``forward`` runs the fused sm_100a quantiser kernel.

Out of scope (training only): ``_update_ema`` / ``_init_ema`` (vq.py:47-94).  ``forward`` in
training mode raises instead of silently running something else.
"""
from __future__ import annotations

from math import prod
from typing import Optional, Tuple

import torch
from torch import nn

from .. import engine as E


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
        self._packed: Optional[E.PackedQuantizer] = None
        self._packed_key = None
        #: number of near-tie vectors (top-2 relative L4 gap < engine.NEAR_TIE_REL_GAP) seen by
        #: the last forward; a 0-dim int32 CUDA tensor (no host sync is forced)
        self.last_near_ties: Optional[torch.Tensor] = None

    # -- packing ---------------------------------------------------------------------------
    def _proj(self):
        return None, None

    def packed(self) -> E.PackedQuantizer:
        proj_in, proj_out = self._proj()
        tensors = [self.embed] + ([proj_in.weight, proj_in.bias, proj_out.weight, proj_out.bias]
                                  if proj_in is not None else [])
        key = tuple((t.data_ptr(), t._version) for t in tensors) + (self.commitment_cost,)
        if self._packed is None or key != self._packed_key:
            self._packed = E.PackedQuantizer(self.embed, self.commitment_cost, proj_in, proj_out)
            self._packed_key = key
        return self._packed

    # -- reference API -----------------------------------------------------------------------
    def embed_code(self, embed_idx: torch.Tensor) -> torch.Tensor:
        """F.embedding(embed_idx, self.embed) (vq.py:44-45): [...] int -> [..., D]."""
        E.require_cuda(embed_idx, "EMAVectorQuantizer.embed_code")
        bare = E.PackedQuantizer(self.embed, self.commitment_cost)  # table == embed
        flat = embed_idx.reshape(-1)
        out = E.embed_codes(bare, flat, True, 1, flat.numel())
        return out.view(*embed_idx.shape, self.embedding_dim)

    def _check(self, inputs: torch.Tensor, channels: int) -> None:
        ndim = inputs.dim()
        assert ndim >= 3                                                    # vq.py:98
        if inputs.shape[1] != channels:                                     # vq.py:100-104
            raise NotImplementedError(
                'VQ dim != channel dim not supported;'
                f' found channel dim of {inputs.shape[1]}, expected {channels}')
        if self.training:
            raise RuntimeError(
                "EMAVectorQuantizer: training-mode forward (EMA codebook update, vq.py:47-94) "
                "is outside the B200 inference path; call .eval()")
        if ndim != 4:
            # the reference passes p = inputs.dim() to cdist (vq.py:121-129); only p = 4 is built
            raise NotImplementedError(
                f"only 4-D inputs (L4 distance) are supported, got {ndim}-D")
        E.require_cuda(inputs, type(self).__name__ + ".forward")

    def _run(self, inputs: torch.Tensor, want_z: bool = False):
        pq = self.packed()
        b = inputs.shape[0]
        s = prod(inputs.shape[2:])
        cl = E.is_channels_last(inputs)
        x = inputs if inputs.dtype == torch.float32 else inputs.float()
        if cl:
            x = x.permute(0, 2, 3, 1).contiguous()      # a view: already NHWC in memory
        else:
            x = x.contiguous()
        out, idx, loss, ties, z = E.quantize(pq, x, cl, cl, b, s, want_out=True, want_z=want_z)
        self.last_near_ties = ties
        sp = tuple(inputs.shape[2:])
        if cl:
            quantized = out.view(b, *sp, pq.c).permute(0, 3, 1, 2)
        else:
            quantized = out.view(b, pq.c, *sp)
        return quantized, idx.view(b, *sp), loss, z

    def forward(self, inputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        self._check(inputs, self.embedding_dim)
        quantized, idx, loss, _ = self._run(inputs)
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

    def _proj(self):
        return self.proj_in, self.proj_out

    def forward(self, inputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        self._check(inputs, self.proj_in.in_channels)
        quantized, idx, loss, _ = self._run(inputs)
        return quantized, idx, loss

    def decode_codes(self, embed_idx: torch.Tensor, channels_last: bool = False) -> torch.Tensor:
        """proj_out(embed_code(idx)) for stored code maps: [B,H,W] int -> [B,C,H,W]."""
        E.require_cuda(embed_idx, "decode_codes")
        pq = self.packed()
        b, h, w = embed_idx.shape
        out = E.embed_codes(pq, embed_idx.reshape(-1), channels_last, b, h * w)
        if channels_last:
            return out.view(b, h, w, pq.c).permute(0, 3, 1, 2)
        return out.view(b, pq.c, h, w)
