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


class BlobNetInputBuffers:
    """Persistent BlobNet / UNet input canvases.

    The reference rebuilds ``blobnet_model_input`` [2B, 4+1+C, h, 2w] and ``unet_bg_input`` [2B, 5, h, 2w] with four
    ``torch.cat`` calls in every denoising step (pipeline_blobnet.py:1043-1049, :1071-1076) although only the 4 noisy
    latent channels of the right half change.  Here both canvases are allocated once; ``fill_static`` writes the
    loop-invariant planes (score maps, the stage-3 feature planes — computed on the fly, ``fg_gs_feats`` is never
    materialised — and the reference-image latents of the left half) with one kernel per canvas, and ``update``
    refreshes the right-half latents per step with two strided copies.  ``blobnet_input`` / ``unet_bg_input`` are then
    bit-identical to ``construct_blobnet_input(...)`` of the reference.
    """

    def __init__(self, batch: int, h: int, w: int, feat_channels: int, dtype: torch.dtype, device="cuda",
                 latent_channels: int = 4):
        self.b, self.h, self.w, self.c, self.lc = batch, h, w, feat_channels, latent_channels
        self.blobnet_input = torch.zeros((batch, latent_channels + 1 + feat_channels, h, 2 * w), dtype=dtype, device=device)
        self.unet_bg_input = torch.zeros((batch, latent_channels + 1, h, 2 * w), dtype=dtype, device=device)

    def fill_static(self, fg_gs_scores: torch.Tensor, bg_gs_scores: torch.Tensor, fg_feats: torch.Tensor,
                    fg_image_latents: torch.Tensor, bg_image_latents: torch.Tensor) -> None:
        """fg/bg_gs_scores [2B,1,h,w]; fg_feats [2B,1,C] (pooled DINOv2 feature, repeated); image latents [2B,4,h,w]."""
        from .. import _capi as C
        dt = self.blobnet_input.dtype
        fg = fg_gs_scores.to(dt).contiguous(); bg = bg_gs_scores.to(dt).contiguous()
        f = fg_feats.to(dt).contiguous()
        k = fg.shape[1]
        C.check(C.lib().blobsplat_conditioning_fill(C.ptr(fg), C.ptr(f), C.ptr(self.blobnet_input), self.b, k, self.c, self.h,
                                                    self.w, self.blobnet_input.shape[1], self.lc, 2, 1, C.dtype_code(dt),
                                                    C.dev_of(fg), C.stream_of(fg)))
        C.check(C.lib().blobsplat_conditioning_fill(C.ptr(bg), None, C.ptr(self.unet_bg_input), self.b, bg.shape[1], 0, self.h,
                                                    self.w, self.unet_bg_input.shape[1], self.lc, 2, 1, C.dtype_code(dt),
                                                    C.dev_of(bg), C.stream_of(bg)))
        self.blobnet_input[:, :self.lc, :, :self.w].copy_(fg_image_latents)      # left half: reference-image latents
        self.unet_bg_input[:, :self.lc, :, :self.w].copy_(bg_image_latents)

    def update(self, latent_model_input: torch.Tensor):
        """Per denoising step: refresh the right-half latents; returns (blobnet_model_input, unet_bg_input)."""
        self.blobnet_input[:, :self.lc, :, self.w:].copy_(latent_model_input)
        self.unet_bg_input[:, :self.lc, :, self.w:].copy_(latent_model_input)
        return self.blobnet_input, self.unet_bg_input


def inject_residual(hidden: torch.Tensor, residual: torch.Tensor, conditioning_scale=1.0) -> torch.Tensor:
    """N4: ``hidden[..., -h:] += conditioning_scale * residual[..., -h:]`` in place, one pass (h = hidden.shape[-2]; on a
    square map the whole width).  Fuses models/blobnet.py:936-938, pipeline_blobnet.py:1085-1087 and
    unet_2d_condition.py:1215-1219.  ``conditioning_scale``: float or a per-sample tensor [B].

    ``residual`` is either the full map [B, C, H, Wr] or the right-hand crop the reference pipeline passes
    (``residual[..., -h:]``, pipeline_blobnet.py:1085-1087 — a strided view): the crop is consumed in place through its
    row stride, never copied."""
    from .. import _capi as C
    C.require_cuda(hidden, "hidden")
    if not hidden.is_contiguous() or hidden.dtype != residual.dtype:
        raise RuntimeError("inject_residual needs a contiguous hidden state and one dtype")
    b, c, h, wh = hidden.shape
    cols = min(h, wh) if wh != h else wh
    rptr, wr = residual.data_ptr(), residual.shape[-1]
    if not residual.is_contiguous():
        # the right-most `cols` columns of rows that are `wr_full` apart: hand the kernel the row start and the row pitch
        sb, sc, sh, sw = residual.stride()
        wr_full = sh
        if (residual.shape[-1] != cols or sw != 1 or sc != h * wr_full or (b > 1 and sb != c * h * wr_full)
                or residual.storage_offset() < wr_full - cols):
            residual = residual.contiguous()
            rptr, wr = residual.data_ptr(), residual.shape[-1]
        else:
            rptr, wr = residual.data_ptr() - (wr_full - cols) * residual.element_size(), wr_full
    if tuple(residual.shape[:3]) != (b, c, h) or wr < cols:
        raise RuntimeError(f"residual {tuple(residual.shape)} does not match hidden {tuple(hidden.shape)}")
    sb = None
    sc = 1.0
    if torch.is_tensor(conditioning_scale):
        sb = conditioning_scale.to(device=hidden.device, dtype=hidden.dtype).float().reshape(b).contiguous()
    else:
        sc = float(conditioning_scale)
    import ctypes
    C.check(C.lib().blobsplat_residual_inject(C.ptr(hidden), ctypes.c_void_p(rptr), C.ptr(sb), sc, b, c, h, wh, wr, cols,
                                              C.dtype_code(hidden.dtype), C.dev_of(hidden), C.stream_of(hidden)))
    return hidden


class BlobConditioningMixin:
    """Method-compatible replacements for StableDiffusionBlobNetPipeline's two conditioning methods."""

    def splat_features_from_scores(self, scores: torch.Tensor, features: torch.Tensor, size: Optional[int],
                                   channels_last: bool = True) -> torch.Tensor:
        return _splat_from_scores(scores, features, size, channels_last=channels_last)

    def construct_blobnet_input(self, latent_model_input: torch.Tensor, gs_scores: torch.Tensor,
                                image_latents: torch.Tensor, gs_feats: torch.Tensor = None,
                                background: bool = False) -> torch.Tensor:
        return construct_blobnet_input(latent_model_input, gs_scores, image_latents, gs_feats, background)
