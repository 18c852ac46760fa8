"""BASELINE configs[3] — the full SD-1.5-shaped UNet + BlobNet edit loop.

Drives the UNMODIFIED reference pipeline class (``StableDiffusionBlobNetPipeline.__call__``,
blobctrl/pipelines/pipeline_blobnet.py:743-1168, installed in baseline/_ref) with random-init weights:

  * UNet  : ``UNet2DConditionModel`` of the reference's diffusers fork, SD-1.5 shapes, conv_in widened 4 -> 5 planes as
            scripts/blobctrl_inference.py:233-249 does (``config.in_channels`` stays 4);
  * BlobNet: ``BlobNetModel(in_channels=4, conditioning_channels=1025, cross_attention_dim=None)`` with the UNet's topology (as
            models/blobnet.py:497-545 / assets/docs/blobnet.txt), its 28 zero-initialised 1x1 taps re-drawn at random so the residual path carries signal;
  * scheduler: UniPC, 50 steps, CFG 7.5 (blobctrl_inference.py:276, :308-311), seed 1248464818;
  * VAE: a small random-init ``AutoencoderKL`` with the SD scale factor 8 (it runs twice per edit, outside the loop);
  * text / DINOv2 encoders: not instantiated — ``prompt_embeds`` [B,77,768] are passed in (a ``__call__`` argument) and
    the pooled DINOv2 feature [1,1,1024] comes from a seeded generator through an overridden ``encode_image_dinov2``
    (:690-703).  Both encoders are outside SURVEY §8 and would be random-init anyway.

Two arms on identical seeded inputs:
  stock  the pipeline exactly as the reference runs it;
  ours   the same object after ``LoopAccelerator(pipe).install()`` (blobctrl_b200/pipelines/loop_accel.py): CUDA stage-3
         splat, persistent canvases (N2), hoisted conv_in (N1), fused residual injection (N4).  ``__call__`` is untouched.
The blob score map fed to both arms comes from the arm's own renderer (reference CPU fp64 recipe vs. the CUDA renderer),
so the comparison covers the whole producer -> loop path.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch

from . import ref_loader

SEED = 1248464818                       # scripts/blobctrl_inference.py:311
ELLIPSE = ((227.1, 118.9), (85.5, 103.7), 87.4)


def sd15_unet_config(small: bool = False) -> dict:
    if small:      # same topology at toy width: CPU-runnable structure test
        return dict(sample_size=16, in_channels=4, out_channels=4, block_out_channels=(32, 64), layers_per_block=1,
                    down_block_types=("CrossAttnDownBlock2D", "DownBlock2D"), up_block_types=("UpBlock2D", "CrossAttnUpBlock2D"),
                    cross_attention_dim=32, attention_head_dim=4, norm_num_groups=8)
    return dict(sample_size=64, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
                up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3, cross_attention_dim=768, attention_head_dim=8)


def build_pipeline(device, dtype=torch.float16, small: bool = False, feat_channels: int = 1024, seed: int = 0):
    """Random-init reference pipeline object (see module docstring)."""
    ns = ref_loader.load()
    torch.manual_seed(seed)
    cfg = sd15_unet_config(small)
    unet = ns.UNet2DConditionModel(**cfg)
    with torch.no_grad():    # scripts/blobctrl_inference.py:233-249: conv_in widened 4 -> 5 planes, config.in_channels stays 4
        wide = torch.nn.Conv2d(cfg["in_channels"] + 1, unet.conv_in.out_channels, kernel_size=3, stride=1, padding=1)
        wide.weight[:, :cfg["in_channels"]].copy_(unet.conv_in.weight)          # (the script zero-fills the new plane; random
        wide.bias.copy_(unet.conv_in.bias)                                      #  here, so the background score plane matters)
        unet.conv_in = wide
    # BlobNet as from_unet builds it (models/blobnet.py:497-545: the UNet's topology, cross_attention_dim=None) but on the 4
    # latent planes: conv_in = Conv2d(4 + 1 + C, 320, 3) as assets/docs/blobnet.txt:2 shows (the script loads it from a
    # checkpoint, blobctrl_inference.py:253; from_unet on the widened 5-plane UNet would give 5 + 1025 input planes)
    topo = {k: cfg[k] for k in ("down_block_types", "up_block_types", "block_out_channels", "layers_per_block", "attention_head_dim")
            if k in cfg}
    if "norm_num_groups" in cfg:
        topo["norm_num_groups"] = cfg["norm_num_groups"]
    blobnet = ns.BlobNetModel(in_channels=4, conditioning_channels=1 + feat_channels, cross_attention_dim=None, **topo)
    with torch.no_grad():
        # from_unet zero-fills the conditioning part of conv_in and the 28 taps (zero_module): re-draw them so the path under
        # test is not multiplied by zero
        w = blobnet.conv_in.weight
        w[:, 4:].normal_(0.0, (9 * w.shape[1]) ** -0.5)
        for tap in list(blobnet.blobnet_down_blocks) + [blobnet.blobnet_mid_block] + list(blobnet.blobnet_up_blocks):
            tap.weight.normal_(0.0, 0.1 * tap.weight.shape[1] ** -0.5)
    vae_ch = (16, 16, 16, 16) if small else (32, 32, 32, 32)
    vae = ns.AutoencoderKL(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=vae_ch, layers_per_block=1,
                           down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
                           norm_num_groups=8, scaling_factor=0.18215)
    sched = ns.UniPCMultistepScheduler()

    class HarnessPipeline(ns.pipeline):
        dino_feature = None

        def encode_image_dinov2(self, image, device):            # stands in for Dinov2Model.pooler_output (:690-703)
            return self.dino_feature.to(device)

    pipe = HarnessPipeline(vae=vae, unet=unet, tokenizer=None, text_encoder=None, blobnet=blobnet, scheduler=sched,
                           safety_checker=None, dinov2_processor=None, dinov2=None, requires_safety_checker=False)
    g = torch.Generator().manual_seed(seed + 1)
    pipe.dino_feature = torch.randn(1, 1, feat_channels, generator=g).to(dtype)
    pipe.unet.to(device=device, dtype=dtype).eval()
    pipe.blobnet.to(device=device, dtype=dtype).eval()
    pipe.vae.to(device=device, dtype=dtype).eval()
    pipe.set_progress_bar_config(disable=True)
    return pipe


def make_inputs(pipe, batch: int, device, dtype, small: bool = False, seed: int = 0) -> dict:
    cfg = pipe.unet.config
    g = torch.Generator().manual_seed(seed + 2)
    hw = 8 * cfg.sample_size
    return {
        "prompt_embeds": torch.randn(batch, 77, cfg.cross_attention_dim, generator=g).to(device=device, dtype=dtype),
        "negative_prompt_embeds": torch.randn(batch, 77, cfg.cross_attention_dim, generator=g).to(device=device, dtype=dtype),
        "fg_image": torch.rand(1, 3, hw, hw, generator=g) * 2 - 1,
        "bg_image": torch.rand(1, 3, hw, hw, generator=g) * 2 - 1,
        "height": hw, "width": hw,
    }


def reference_gs_score(size: int) -> torch.Tensor:
    """The producer as the reference's scripts run it: CPU, float64 (scripts/blobctrl_inference.py:71-117, :170-174)."""
    import numpy as np
    U = ref_loader.load(need_pipeline=False).utils
    (xc, yc), (d1, d2), ang = ELLIPSE
    theta = np.radians((((180 - ang) % 180) + 90) % 180)
    mean, cov = U.ellipse_to_gaussian(xc, yc, d1 / 2, d2 / 2, theta)
    nm, nc = mean / np.array([512, 512]), cov / (512 ** 2 + 512 ** 2)
    blob = {"xs": torch.tensor(nm[0]).unsqueeze(0), "ys": torch.tensor(nm[1]).unsqueeze(0),
            "covs": torch.tensor(nc).unsqueeze(0).unsqueeze(0), "sizes": torch.tensor([1.0]).unsqueeze(0)}
    return U.splat_features(**blob, score_size=(size, size), return_d_score=True)[0].unsqueeze(0)     # [1,2,s,s] fp64 CPU


def cuda_gs_score(size: int, device) -> torch.Tensor:
    """The same map from the CUDA renderer's ellipse front end (one launch, on the device)."""
    import blobctrl_b200 as B
    e = torch.tensor([[list(ELLIPSE[0]) + list(ELLIPSE[1]) + [ELLIPSE[2]]]], dtype=torch.float32, device=device)
    return B.splat_ellipses(e, image_size=(512, 512), score_size=size)                                # [1,2,s,s] fp32 CUDA


@torch.no_grad()
def run_edit(pipe, inputs: dict, gs_score: torch.Tensor, steps: int, device, autocast_dtype: Optional[torch.dtype]):
    """One ``pipe(...)`` call; returns (latents [B,4,h,w], seconds by CUDA events or wall clock on CPU)."""
    cuda = torch.device(device).type == "cuda"
    torch.manual_seed(SEED)
    gen = torch.Generator(device=device).manual_seed(SEED)
    kw = dict(prompt=None, fg_image=inputs["fg_image"], bg_image=inputs["bg_image"], gs_score=gs_score, height=inputs["height"],
              width=inputs["width"], num_inference_steps=steps, guidance_scale=7.5, generator=gen, output_type="latent",
              prompt_embeds=inputs["prompt_embeds"], negative_prompt_embeds=inputs["negative_prompt_embeds"], return_dict=False)
    if cuda:
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    t0 = time.perf_counter()
    if autocast_dtype is not None and cuda:
        with torch.autocast("cuda", dtype=autocast_dtype):
            out = pipe(**kw)[0]
    else:
        out = pipe(**kw)[0]
    if cuda:
        e1.record(); torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) / 1e3
    return out, time.perf_counter() - t0


def compare_arms(device="cuda", dtype=torch.float16, batch: int = 8, steps: int = 50, small: bool = False, warm_steps: int = 2,
                 variants=(("ours_n2_n4", dict(hoist_conv_in=False)), ("ours_n1_n2_n4", dict(hoist_conv_in=True)))) -> Dict:
    """Stock loop vs. accelerated loop on identical inputs.  Returns timings (s/edit), latent differences, the splat's
    measured share and the accelerator's call counters."""
    from blobctrl_b200.pipelines.loop_accel import LoopAccelerator
    pipe = build_pipeline(device, dtype, small)
    size = pipe.unet.config.sample_size
    inputs = make_inputs(pipe, batch, device, dtype, small)
    ac = dtype if dtype != torch.float32 else None
    res = {"batch": batch, "cfg_batch": 2 * batch, "steps": steps, "dtype": str(dtype).replace("torch.", ""), "latent": size,
           "unet_params_m": sum(p.numel() for p in pipe.unet.parameters()) / 1e6,
           "blobnet_params_m": sum(p.numel() for p in pipe.blobnet.parameters()) / 1e6}

    t0 = time.perf_counter(); gs_ref = reference_gs_score(size); res["producer_reference_cpu_ms"] = (time.perf_counter() - t0) * 1e3
    run_edit(pipe, inputs, gs_ref, warm_steps, device, ac)                               # warm-up (cuDNN heuristics, allocator)
    lat_stock, res["stock_s_per_edit"] = run_edit(pipe, inputs, gs_ref, steps, device, ac)
    scale = float(lat_stock.float().abs().max())
    res["latent_absmax"] = scale

    cuda_gs_score(size, device); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gs_cuda = cuda_gs_score(size, device); e1.record(); torch.cuda.synchronize()
    res["producer_cuda_ms"] = e0.elapsed_time(e1)
    res["gs_score_max_abs_diff"] = float((gs_cuda.double().cpu() - gs_ref).abs().max())
    for name, opts in variants:
        acc = LoopAccelerator(pipe, **opts).install()
        try:
            run_edit(pipe, inputs, gs_cuda, warm_steps, device, ac)
            for k in acc.stats:
                acc.stats[k] = 0
            lat, sec = run_edit(pipe, inputs, gs_cuda, steps, device, ac)
        finally:
            acc.remove()
        diff = (lat.float() - lat_stock.float()).abs()
        res[name] = {"s_per_edit": sec, "speedup_vs_stock": res["stock_s_per_edit"] / sec,
                     "latent_max_abs_diff": float(diff.max()), "latent_rel_to_absmax": float(diff.max()) / max(scale, 1e-30),
                     "latent_rms_diff": float(diff.pow(2).mean().sqrt()), "bit_identical": bool(torch.equal(lat, lat_stock)),
                     "calls": dict(acc.stats)}
    # the splat's share of an edit: the producer + the once-per-edit stage 3, against the stock edit
    import blobctrl_b200 as B
    fg = gs_cuda[:, 1:2].repeat(2 * batch, 1, 1, 1).to(dtype)
    f = pipe.dino_feature.to(device).repeat(2 * batch, 1, 1)
    B.splat_features_from_scores(fg, f, size, channels_last=False); torch.cuda.synchronize()
    e0.record(); B.splat_features_from_scores(fg, f, size, channels_last=False); e1.record(); torch.cuda.synchronize()
    res["stage3_cuda_ms"] = e0.elapsed_time(e1)
    res["splat_share_of_stock_edit"] = (res["producer_cuda_ms"] + res["stage3_cuda_ms"]) / 1e3 / res["stock_s_per_edit"]
    return res


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--dtype", default="float16")
    a = ap.parse_args()
    print(json.dumps(compare_arms(batch=a.batch, steps=a.steps, small=a.small, dtype=getattr(torch, a.dtype))))
