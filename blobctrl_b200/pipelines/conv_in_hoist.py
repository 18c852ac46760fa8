"""N1: loop-invariant hoist of BlobNet's ``conv_in``.

BlobNet's first layer is ``Conv2d(4 + 1 + C, 320, 3, padding=1)`` over the canvas built by
``construct_blobnet_input`` (blobctrl/models/blobnet.py:241-245, applied at :840; input built at
blobctrl/pipelines/pipeline_blobnet.py:1043-1049).  Of its 4+1+C input channels only the 4 latent channels of the
right half change between the 50 denoising steps, and the C feature planes are rank-K: feats[c] = sum_k s_k * f[k, c].
Convolution is linear in its input, so

    conv_in(x) = bias + conv(W[:, :4], latent canvas)                         <- changes per step
               + conv(W[:, 4:5], score) + sum_k conv(W_eff[:, k], s_k)        <- loop-invariant planes
    W_eff[b, o, k] = sum_c W[o, 5 + c] * f[b, k, c]      (K effective 3x3 kernels instead of C = 1024 input planes)

which removes a 1029 -> 320 3x3 convolution (~48.6 GFLOP per sample per step) and never materialises the 1024
feature planes or the 1029-plane canvas.  Both pieces are hand-written CUDA behind the C ABI
(``blobsplat_conv_in_weights`` once per edit, ``blobsplat_conv_in_hoisted`` per step — a (4 + 1 + K) -> 320 direct
convolution with per-sample kernels for the conditioning planes, fp32 accumulation, one rounding).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops


class HoistedConvIn(torch.nn.Module):
    """Drop-in for ``blobnet.conv_in`` inside the denoising loop.

    weight [O, 4+1+C, 3, 3], bias [O]: BlobNet's conv_in parameters (shared, not copied).  ``prepare`` is called once
    per edit with the loop-invariant conditioning; the module is then called once per step with the 4-channel latent
    canvas [2B, 4, h, 2w] (left = reference-image latents, right = noisy latents, as in construct_blobnet_input).
    A full (4+1+C)-plane canvas is still accepted and goes through the original convolution.
    """

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], latent_channels: int = 4, halves: int = 2):
        super().__init__()
        self.lc, self.halves = latent_channels, halves
        self.weight, self.bias = weight, bias
        self.cond: Optional[torch.Tensor] = None
        self.weff: Optional[torch.Tensor] = None
        self._cast = None

    @classmethod
    def from_conv(cls, conv: torch.nn.Conv2d, latent_channels: int = 4, halves: int = 2) -> "HoistedConvIn":
        if conv.kernel_size != (3, 3) or conv.padding != (1, 1) or conv.stride != (1, 1):
            raise RuntimeError("HoistedConvIn replaces a 3x3, stride 1, padding 1 convolution")
        return cls(conv.weight.detach(), None if conv.bias is None else conv.bias.detach(), latent_channels, halves)

    def _compute_dtype(self) -> torch.dtype:
        """The dtype the reference layer computes in: autocast's when it is on (blobctrl_inference.py:190), else the weights'."""
        if torch.is_autocast_enabled():
            return torch.get_autocast_dtype("cuda")
        return self.weight.dtype

    def _params(self, dt: torch.dtype):
        if self._cast is None or self._cast[0] != dt:
            self._cast = (dt, self.weight.to(dt).contiguous(), None if self.bias is None else self.bias.to(dt).contiguous())
        return self._cast[1], self._cast[2]

    @torch.no_grad()
    def prepare(self, gs_scores: torch.Tensor, blob_scores: torch.Tensor, feats: torch.Tensor) -> None:
        """gs_scores [2B,1,h,w]: the score plane of the canvas; blob_scores [2B,K,h,w] and feats [2B,K,C]: the
        stage-3 operands (in the reference pipeline K = 1 and blob_scores is gs_scores, pipeline_blobnet.py:984)."""
        dt = self._compute_dtype()
        w, _ = self._params(dt)
        self.cond = torch.cat([gs_scores, blob_scores], dim=1).to(dt).contiguous()       # [2B, 1+K, h, w]
        self.weff = ops.conv_in_weights(w, feats, self.lc)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dt = self._compute_dtype()
        w, b = self._params(dt)
        if x.shape[1] != self.lc:                                   # a full canvas: the layer as the reference runs it
            return torch.nn.functional.conv2d(x.to(dt), w, b, padding=1)
        if self.weff is None:
            raise RuntimeError("HoistedConvIn.prepare() has not been called for this edit")
        return ops.conv_in_hoisted(x.to(dt).contiguous(), self.cond.to(dt), w, b, self.weff, halves=self.halves)
