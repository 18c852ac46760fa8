"""N1: loop-invariant hoist of BlobNet's ``conv_in``.

BlobNet's first layer is ``Conv2d(4 + 1 + C, 320, 3, padding=1)`` over the canvas built by
``construct_blobnet_input`` (blobctrl/models/blobnet.py:241-245, applied at :840; input built at
blobctrl/pipelines/pipeline_blobnet.py:1043-1049).  Of its 4+1+C input channels only the 4 latent channels of the
right half change between the 50 denoising steps, and the C feature planes are rank-K: feats[c] = sum_k s_k * f[k, c].
Convolution is linear in its input, so

    conv_in(x) = conv(W[:, :4], latents_canvas)                               <- per step, 4 -> 320 channels
               + conv(W[:, 4:5], scores) + conv(W_eff, scores_k) + bias       <- once per edit
    W_eff[o, k] = sum_c W[o, 5 + c] * f[k, c]      (K effective 3x3 kernels instead of C = 1024 input planes)

which removes a 1029 -> 320 3x3 convolution (~48.6 GFLOP per sample per step) and never materialises the 1024
feature planes at all.  The convolutions themselves stay library calls (cuDNN via torch) — they are outside the splat
hot path; what this module contributes is the algebra and the exactness check (tests: fp32 1e-4 of scale).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class HoistedConvIn:
    """Drop-in for ``blobnet.conv_in`` inside the denoising loop.

    weight [O, 4+1+C, 3, 3], bias [O]: BlobNet's conv_in parameters.  ``prepare`` is called once per edit with the
    loop-invariant conditioning; ``__call__`` once per step with the 4-channel latent canvas [2B, 4, h, 2w]
    (left = reference-image latents, right = noisy latents, as in construct_blobnet_input).
    """

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, latent_channels: int = 4):
        self.lc = latent_channels
        self.w_lat = weight[:, :latent_channels].contiguous()
        self.w_score = weight[:, latent_channels:latent_channels + 1].contiguous()
        self.w_feat = weight[:, latent_channels + 1:].contiguous()          # [O, C, 3, 3]
        self.bias = bias
        self.static = None

    @torch.no_grad()
    def prepare(self, gs_scores: torch.Tensor, blob_scores: torch.Tensor, feats: torch.Tensor) -> torch.Tensor:
        """gs_scores [2B,1,h,w]: the score plane of the canvas; blob_scores [2B,K,h,w] and feats [2B,K,C]: the
        stage-3 operands (in the reference pipeline K = 1 and blob_scores is gs_scores).  Both halves of the canvas
        carry the same conditioning, so the static part is computed on the width-doubled maps."""
        two = lambda t: torch.cat([t, t], dim=-1)
        # W_eff[b, o, k, :, :] = sum_c W[o, c] * f[b, k, c]  -> per-sample grouped conv with K input planes
        w_eff = torch.einsum("ocij,bkc->bokij", self.w_feat.float(), feats.float()).to(self.w_lat.dtype)
        b, o, k = w_eff.shape[:3]
        s2 = two(blob_scores)                                               # [2B, K, h, 2w]
        feat_part = F.conv2d(s2.reshape(1, b * k, *s2.shape[-2:]), w_eff.reshape(b * o, k, 3, 3), padding=1, groups=b)
        feat_part = feat_part.reshape(b, o, *s2.shape[-2:])
        self.static = F.conv2d(two(gs_scores), self.w_score, self.bias, padding=1) + feat_part
        return self.static

    @torch.no_grad()
    def __call__(self, latent_canvas: torch.Tensor) -> torch.Tensor:
        return F.conv2d(latent_canvas, self.w_lat, None, padding=1) + self.static
