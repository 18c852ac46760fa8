"""BlobNet's conditioning entry — the prologue of ``StableDiffusionBlobNetPipeline.__call__``.

Mirrors, with unchanged call signatures and tensor layouts, the three pieces of
``blobctrl/pipelines/pipeline_blobnet.py`` that touch the blob maps:

  * ``splat_features_from_scores``  (:706-721)  — stage 3 on the CUDA kernels;
  * ``construct_blobnet_input``     (:724-739)  — channel order (latent, score, feats), left half =
    reference-image latents, right half = noisy latents, concatenated along the width;
  * the prologue at :973-984 — unbind bg/fg scores, repeat to the CFG batch, splat the pooled DINOv2
    feature (K = 1, C = 1024) into ``fg_gs_feats``.

``BlobConditioningMixin`` can be mixed into (or monkey-patched onto) the reference pipeline class:
its two methods have the reference's names and signatures, so the 50-step loop is untouched.
The rest of the pipeline (VAE, UNet, BlobNet, scheduler) is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from ..utils.utils import splat_features_from_scores as _splat_from_scores


def splat_features_from_scores(scores: torch.Tensor, features: torch.Tensor, size: Optional[int],
                               channels_last: bool = True) -> torch.Tensor:
    """pipeline_blobnet.py:706-721 (function form)."""
    return _splat_from_scores(scores, features, size, channels_last=channels_last)


def construct_blobnet_input(latent_model_input: torch.Tensor, gs_scores: torch.Tensor, image_latents: torch.Tensor,
                            gs_feats: torch.Tensor = None, background: bool = False) -> torch.Tensor:
    """pipeline_blobnet.py:724-739.  Returns [2B, 4+1(+C), h, 2w]: left = image latents, right = noisy
    latents, each followed by the score map and (foreground only) the feature map."""
    if not background:
        right = torch.cat([latent_model_input, gs_scores, gs_feats], dim=1)
        left = torch.cat([image_latents, gs_scores, gs_feats], dim=1)
    else:
        right = torch.cat([latent_model_input, gs_scores], dim=1)
        left = torch.cat([image_latents, gs_scores], dim=1)
    return torch.cat([left, right], dim=-1)


@dataclass
class BlobConditioning:
    """Loop-invariant conditioning tensors (pipeline_blobnet.py:974-984)."""
    bg_gs_scores: torch.Tensor   # [2B, 1, h, w]
    fg_gs_scores: torch.Tensor   # [2B, 1, h, w]
    fg_gs_feats: torch.Tensor    # [2B, C, h, w]


def prepare_blob_conditioning(gs_score: torch.Tensor, fg_feats: torch.Tensor, batch: int,
                              dtype: torch.dtype, device) -> BlobConditioning:
    """pipeline_blobnet.py:973-984.

    gs_score: [1, 2, h, w] (bg, fg) as returned by ``splat_features(..., return_d_score=True)``;
    fg_feats: [1, 1, C] pooled image feature (DINOv2 pooler_output, :690-703);
    batch: ``prompt_embeds.shape[0]`` (2B with classifier-free guidance).
    """
    bg, fg = gs_score.unbind(dim=1)
    if bg.ndim == 3:
        bg, fg = bg.unsqueeze(1), fg.unsqueeze(1)
    bg = bg.repeat(batch, 1, 1, 1).to(device=device, dtype=dtype)
    fg = fg.repeat(batch, 1, 1, 1).to(device=device, dtype=dtype)
    feats = fg_feats.to(device=device).repeat(batch, 1, 1)
    fg_gs_feats = _splat_from_scores(fg, feats, size=fg.shape[2], channels_last=False)
    return BlobConditioning(bg, fg, fg_gs_feats)


class BlobConditioningMixin:
    """Method-compatible replacements for StableDiffusionBlobNetPipeline's two conditioning methods."""

    def splat_features_from_scores(self, scores: torch.Tensor, features: torch.Tensor, size: Optional[int],
                                   channels_last: bool = True) -> torch.Tensor:
        return _splat_from_scores(scores, features, size, channels_last=channels_last)

    def construct_blobnet_input(self, latent_model_input: torch.Tensor, gs_scores: torch.Tensor,
                                image_latents: torch.Tensor, gs_feats: torch.Tensor = None,
                                background: bool = False) -> torch.Tensor:
        return construct_blobnet_input(latent_model_input, gs_scores, image_latents, gs_feats, background)
