"""Drop-in acceleration of the reference's editing loop — N1 / N2 / N4 of SURVEY.md §8(f) inside the loop they exist for.

``LoopAccelerator(pipe).install()`` takes a constructed ``StableDiffusionBlobNetPipeline`` (the UNMODIFIED reference
class, blobctrl/pipelines/pipeline_blobnet.py) and substitutes, without editing a line of the reference or of its
diffusers fork:

  * ``pipe.splat_features_from_scores``   (:706-721, called once per edit at :984)  -> the CUDA stage-3 kernels;
  * ``pipe.construct_blobnet_input``      (:724-739, called twice per step at :1043-1049 and :1071-1076) -> persistent
    canvases whose loop-invariant planes are written once per edit (N2); per step only the 4 noisy-latent planes of the
    right half are refreshed;
  * ``pipe.blobnet.conv_in``              (models/blobnet.py:241-245, :840) -> ``HoistedConvIn`` (N1): the canvas
    shrinks to its 4 latent planes and the 1029 -> 320 convolution to a (4 + 1 + K) -> 320 one;
  * the 28 residual additions inside the UNet fork (unet_2d_condition.py:1215-1219, :1292-1296; unet_2d_blocks.py
    :1303-1319, :1411-1428, :2598-2624, :2740-2765) -> forward hooks that apply ``blobsplat_residual_inject`` in place
    (N4); the UNet is then called without ``*_add_samples`` so the fork's own additions do not run.

``pipe.__call__`` — the 50-step loop itself (:1024-1123) — stays the reference's code.  ``remove()`` restores everything.
The N4 hooks apply only to side-by-side canvases (width != height), where the fork's addition is in place too; on a
square canvas the fork's out-of-place ``sample = sample + residual`` differs in what the skip connections see, so the
original path is kept there.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import _capi as C
from .. import ops
from .conditioning import inject_residual
from .conv_in_hoist import HoistedConvIn


class LoopAccelerator:
    def __init__(self, pipe, hoist_conv_in: bool = True, fuse_injection: bool = True, persistent_inputs: bool = True,
                 latent_channels: int = 4):
        self.pipe = pipe
        self.hoist, self.fuse, self.persist, self.lc = hoist_conv_in, fuse_injection, persistent_inputs, latent_channels
        self._installed = False
        self._saved = {}
        self._hooks: List = []
        self._queue: Optional[List[torch.Tensor]] = None
        self._fg_feats = None                      # [2B, K, C] recorded at :984
        self._edit = 0                             # bumped once per pipeline call (a new edit invalidates the canvases)
        self._canvas = {}                          # background flag -> (edit id, tensor)
        self.hoisted: Optional[HoistedConvIn] = None
        self.stats = {"splat_calls": 0, "canvas_fills": 0, "canvas_updates": 0, "injections": 0}

    # ---- substitutes for the two pipeline methods ------------------------------------------------------------------
    def splat_features_from_scores(self, scores, features, size, channels_last=True):
        """pipeline_blobnet.py:706-721 on the CUDA kernels; also the once-per-edit hook point (:984)."""
        from ..utils.utils import splat_features_from_scores as splat
        self._edit += 1
        self._fg_feats = features
        self.stats["splat_calls"] += 1
        if self.hoist and self.hoisted is not None:
            # the feature planes are never read again (the hoisted conv_in consumes scores and features directly): the
            # loop only passes this tensor back into construct_blobnet_input, so a view of the scores stands in for it
            return scores
        return splat(scores, features, size, channels_last=channels_last)

    def construct_blobnet_input(self, latent_model_input, gs_scores, image_latents, gs_feats=None, background=False):
        """pipeline_blobnet.py:724-739 with the loop-invariant planes written once per edit."""
        if not self.persist:
            from .conditioning import construct_blobnet_input
            return construct_blobnet_input(latent_model_input, gs_scores, image_latents, gs_feats, background)
        lc = self.lc
        b, _, h, w = latent_model_input.shape
        dt, dev = latent_model_input.dtype, latent_model_input.device
        hoisted = self.hoist and self.hoisted is not None and not background
        planes = lc if hoisted else (lc + 1 if background else lc + 1 + self._fg_feats.shape[-1])
        key = bool(background)
        ent = self._canvas.get(key)
        if ent is None or ent[0] != self._edit or ent[1].shape != (b, planes, h, 2 * w) or ent[1].dtype != dt:
            buf = ent[1] if ent is not None and ent[1].shape == (b, planes, h, 2 * w) and ent[1].dtype == dt else \
                torch.empty((b, planes, h, 2 * w), dtype=dt, device=dev)
            buf[:, :lc, :, :w].copy_(image_latents)                                  # left half: reference-image latents
            if hoisted:
                self.hoisted.prepare(gs_scores, gs_scores, self._fg_feats)
            else:
                sc = gs_scores.to(dt).contiguous()
                f = None if background else self._fg_feats.to(dt).contiguous()
                C.check(C.lib().blobsplat_conditioning_fill(C.ptr(sc), C.ptr(f), C.ptr(buf), b, sc.shape[1], 0 if f is None else f.shape[-1],
                                                            h, w, planes, lc, 2, 1, C.dtype_code(dt), C.dev_of(buf), C.stream_of(buf)))
            self._canvas[key] = ent = (self._edit, buf)
            self.stats["canvas_fills"] += 1
        ent[1][:, :lc, :, w:].copy_(latent_model_input)                              # right half: this step's noisy latents
        self.stats["canvas_updates"] += 1
        return ent[1]

    # ---- N4: residual injection through forward hooks -----------------------------------------------------------------
    def _injection_sites(self, unet) -> List[torch.nn.Module]:
        """The modules after whose output the fork adds a BlobNet residual, in the order the fork pops them."""
        sites = [unet.conv_in]                                                       # unet_2d_condition.py:1215-1219
        for blk in unet.down_blocks:                                                 # unet_2d_blocks.py:1303-1319, 1411-1428
            attns = getattr(blk, "attentions", None)
            for i in range(len(blk.resnets)):
                sites.append(attns[i] if attns is not None and getattr(blk, "has_cross_attention", False) else blk.resnets[i])
            if blk.downsamplers is not None:
                sites.extend(blk.downsamplers)
        sites.append(unet.mid_block)                                                 # unet_2d_condition.py:1292-1296
        for blk in unet.up_blocks:                                                   # unet_2d_blocks.py:2598-2624, 2740-2765
            attns = getattr(blk, "attentions", None)
            for i in range(len(blk.resnets)):
                sites.append(attns[i] if attns is not None and getattr(blk, "has_cross_attention", False) else blk.resnets[i])
            if blk.upsamplers is not None:
                sites.extend(blk.upsamplers)
        return sites

    def _make_hook(self, index: int):
        def hook(module, args, output):
            q = self._queue
            if q is None:
                return None
            t = output[0] if isinstance(output, tuple) else getattr(output, "sample", output)
            inject_residual(t, q[index], 1.0)           # BlobNet has already applied conditioning_scale (blobnet.py:936-938)
            self.stats["injections"] += 1
            return None
        return hook

    def _unet_forward(self, sample, timestep, encoder_hidden_states, *args, down_block_add_samples=None,
                      mid_block_add_sample=None, up_block_add_samples=None, **kwargs):
        fwd = self._saved["unet_forward"]
        blobnet = down_block_add_samples is not None and mid_block_add_sample is not None and up_block_add_samples is not None
        if not (self.fuse and blobnet and sample.shape[-1] != sample.shape[-2]):
            return fwd(sample, timestep, encoder_hidden_states, *args, down_block_add_samples=down_block_add_samples,
                       mid_block_add_sample=mid_block_add_sample, up_block_add_samples=up_block_add_samples, **kwargs)
        queue = list(down_block_add_samples) + [mid_block_add_sample] + list(up_block_add_samples)
        if len(queue) != len(self._hooks):
            raise RuntimeError(f"{len(queue)} BlobNet residuals for {len(self._hooks)} injection sites")
        self._queue = queue
        try:
            return fwd(sample, timestep, encoder_hidden_states, *args, **kwargs)
        finally:
            self._queue = None

    # ---- install / remove ---------------------------------------------------------------------------------------------
    def install(self) -> "LoopAccelerator":
        if self._installed:
            return self
        pipe = self.pipe
        self._saved = {"splat": pipe.__dict__.get("splat_features_from_scores"),
                       "construct": pipe.__dict__.get("construct_blobnet_input")}
        pipe.splat_features_from_scores = self.splat_features_from_scores
        pipe.construct_blobnet_input = self.construct_blobnet_input
        if self.hoist:
            self._saved["conv_in"] = pipe.blobnet.conv_in
            self.hoisted = HoistedConvIn.from_conv(pipe.blobnet.conv_in, self.lc, halves=2)
            pipe.blobnet.conv_in = self.hoisted
        if self.fuse:
            unet = pipe.unet
            self._saved["unet_forward"] = unet.forward
            for i, m in enumerate(self._injection_sites(unet)):
                self._hooks.append(m.register_forward_hook(self._make_hook(i)))
            unet.forward = self._unet_forward
        self._installed = True
        return self

    def remove(self) -> None:
        if not self._installed:
            return
        pipe = self.pipe
        for name, key in (("splat_features_from_scores", "splat"), ("construct_blobnet_input", "construct")):
            if self._saved.get(key) is None:
                pipe.__dict__.pop(name, None)
            else:
                setattr(pipe, name, self._saved[key])
        if "conv_in" in self._saved:
            pipe.blobnet.conv_in = self._saved["conv_in"]
            self.hoisted = None
        if "unet_forward" in self._saved:
            pipe.unet.__dict__.pop("forward", None)
            for h in self._hooks:
                h.remove()
            self._hooks = []
        self._canvas.clear()
        self._installed = False
