#!/usr/bin/env python
"""Golden fixtures for the pipeline's two conditioning methods, recorded from the UNMODIFIED reference class.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_pipeline.py      ->  tests/golden/pipeline.npz + pipeline_cases.json

Records ``StableDiffusionBlobNetPipeline.splat_features_from_scores`` (blobctrl/pipelines/pipeline_blobnet.py:706-721)
and ``.construct_blobnet_input`` (:724-739) — called unbound, neither touches ``self`` — on CPU in float16 and float32,
plus the conditioning prologue (:973-984) driven through those methods with the repeat / cast order of the reference.
Kept separate from golden.npz so that file stays byte-identical to what make_golden.py writes.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/diffusers/src"); sys.path.insert(0, "/root/reference")
import transformers.utils as TU  # noqa: E402
if not hasattr(TU, "FLAX_WEIGHTS_NAME"):
    TU.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
import diffusers  # noqa: E402,F401  (before the stub: its find_spec("matplotlib") probe)
_m = types.ModuleType("matplotlib"); _m.cm = types.ModuleType("matplotlib.cm")
sys.modules.setdefault("matplotlib", _m); sys.modules.setdefault("matplotlib.cm", _m.cm)
from blobctrl.pipelines.pipeline_blobnet import StableDiffusionBlobNetPipeline as P  # noqa: E402
import blobctrl.utils.utils as REF  # noqa: E402

torch.set_num_threads(1)
arrays, cases = {}, []
g = torch.Generator().manual_seed(17)


def put(key, t):
    # float16 has a numpy dtype; bfloat16 does not and is not used by the reference pipeline
    arrays[key] = t.detach().cpu().numpy()


# ---- splat_features_from_scores (method, :706-721) ---------------------------------------------------
for name, (n, k, h, w, c, size, cl, dt, stride) in {
        "m_s3_k1_f16": (2, 1, 64, 64, 1024, 64, False, torch.float16, 128),     # the pipeline's own call (:984)
        "m_s3_k1_f32": (2, 1, 64, 64, 1024, 64, False, torch.float32, 128),
        "m_s3_k3_f16": (2, 3, 16, 16, 24, 16, False, torch.float16, 1),
        "m_s3_cl_f32": (2, 5, 12, 12, 9, 12, True, torch.float32, 1),
        "m_s3_resize_f32": (1, 4, 24, 24, 6, 16, False, torch.float32, 1),      # bilinear branch (:714-717)
        "m_s3_resize_cl_f32": (1, 3, 16, 16, 4, 40, True, torch.float32, 1)}.items():
    sc = torch.rand((n, h, w, k) if cl else (n, k, h, w), generator=g).to(dt)
    ft = torch.randn((n, k, c), generator=g)                                   # features arrive in another dtype (:713)
    out = P.splat_features_from_scores(None, sc, ft, size, channels_last=cl)
    assert out.dtype == dt and out.is_contiguous()
    assert torch.equal(out, REF.splat_features_from_scores(sc, ft, size, channels_last=cl))   # the duplicate in utils.py:57-77
    put(f"{name}/scores", sc); put(f"{name}/features", ft); put(f"{name}/out", out[:, ::stride])
    cases.append({"name": name, "func": "splat_features_from_scores", "size": size, "channels_last": cl, "c_stride": stride})

# ---- construct_blobnet_input (method, :724-739) ------------------------------------------------------
for name, (b, h, w, c, dt) in {"m_cat_f16": (2, 8, 8, 6, torch.float16), "m_cat_f32": (3, 6, 10, 5, torch.float32)}.items():
    lat = torch.randn(b, 4, h, w, generator=g).to(dt); img = torch.randn(b, 4, h, w, generator=g).to(dt)
    sc = torch.rand(b, 1, h, w, generator=g).to(dt); ft = torch.randn(b, c, h, w, generator=g).to(dt)
    fg = P.construct_blobnet_input(None, lat, sc, img, ft, background=False)
    bg = P.construct_blobnet_input(None, lat, sc, img, background=True)
    assert fg.shape == (b, 5 + c, h, 2 * w) and bg.shape == (b, 5, h, 2 * w)
    for k, v in (("lat", lat), ("img", img), ("scores", sc), ("feats", ft), ("fg", fg), ("bg", bg)):
        put(f"{name}/{k}", v)
    cases.append({"name": name, "func": "construct_blobnet_input"})

# ---- the conditioning prologue (:973-984) and one loop step's canvases (:1043-1049, :1071-1076) --------
for name, dt in (("m_prologue_f16", torch.float16), ("m_prologue_f32", torch.float32)):
    batch = 4                                                                   # prompt_embeds.shape[0] with CFG
    ell = ((227.1, 118.9), (85.5, 103.7), 87.4)
    (xc, yc), (d1, d2), ang = ell
    mean, cov = REF.ellipse_to_gaussian(xc, yc, d1 / 2, d2 / 2, np.radians((((180 - ang) % 180) + 90) % 180))
    nm, nc = mean / np.array([512, 512]), cov / (512 ** 2 + 512 ** 2)
    blob = {"xs": torch.tensor(nm[0]).unsqueeze(0), "ys": torch.tensor(nm[1]).unsqueeze(0),
            "covs": torch.tensor(nc).unsqueeze(0).unsqueeze(0), "sizes": torch.tensor([1.0]).unsqueeze(0)}
    gs_score = REF.splat_features(**blob, score_size=(32, 32), return_d_score=True)[0].unsqueeze(0)   # [1,2,32,32] fp64
    dino = torch.randn(1, 1, 64, generator=g).to(dt)
    bg_s, fg_s = gs_score.unbind(dim=1)                                         # :974
    bg_s = bg_s.unsqueeze(1).repeat(batch, 1, 1, 1).to(dtype=dt)                # :975-979
    fg_s = fg_s.unsqueeze(1).repeat(batch, 1, 1, 1).to(dtype=dt)
    feats = P.splat_features_from_scores(None, fg_s, dino.repeat(batch, 1, 1), size=fg_s.shape[2], channels_last=False)   # :983-984
    lat = torch.randn(batch, 4, 32, 32, generator=g).to(dt)
    fg_lat = torch.randn(batch, 4, 32, 32, generator=g).to(dt); bg_lat = torch.randn(batch, 4, 32, 32, generator=g).to(dt)
    x = P.construct_blobnet_input(None, lat, fg_s, fg_lat, feats, background=False)
    xb = P.construct_blobnet_input(None, lat, bg_s, bg_lat, background=True)
    for k, v in (("gs_score", gs_score), ("dino", dino), ("lat", lat), ("fg_lat", fg_lat), ("bg_lat", bg_lat),
                 ("fg_gs_feats", feats), ("blobnet_model_input", x), ("unet_bg_input", xb)):
        put(f"{name}/{k}", v)
    cases.append({"name": name, "func": "prologue", "batch": batch})

np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **arrays)
json.dump({"torch": torch.__version__, "cases": cases}, open(os.path.join(HERE, "pipeline_cases.json"), "w"), indent=1)
print(f"{len(cases)} cases, {len(arrays)} arrays, {os.path.getsize(os.path.join(HERE, 'pipeline.npz')) / 1e6:.2f} MB")
