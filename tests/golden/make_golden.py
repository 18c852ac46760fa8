#!/usr/bin/env python
"""Generate the golden fixtures by executing the UNMODIFIED reference renderer.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

It imports ``/root/reference/blobctrl/utils/utils.py`` (with a sys.modules stub for the unused
``matplotlib`` import at utils.py:11), runs it on CPU on the inputs below and writes
``tests/golden/golden.npz`` + ``tests/golden/cases.json``.  While doing so it also asserts that
``oracle/aten_port.py`` is bit-identical to the reference on every case (same ATen kernels).

The reference ships no tests or golden vectors; these fixtures are what pins
parity.  Inputs: the 40 real ellipses of ``assets/results/demo/*/state/state.json`` through the
script recipe (``scripts/blobctrl_inference.py:71-117``), the UI preview call
(``scripts/blobctrl_app.py:637-646``), seeded synthetic [N, M] sets and edge cases.
"""
import glob
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
_m = types.ModuleType("matplotlib"); _m.cm = types.ModuleType("matplotlib.cm")
sys.modules.setdefault("matplotlib", _m); sys.modules.setdefault("matplotlib.cm", _m.cm)

import blobctrl.utils.utils as REF  # noqa: E402  (the reference, unmodified)

from oracle import aten_port, blob_oracle  # noqa: E402

torch.set_num_threads(1)
arrays, cases = {}, []


def put(name, kind, key, val):
    arrays[f"{name}/{kind}/{key}"] = val.detach().cpu().numpy() if torch.is_tensor(val) else np.asarray(val)


def flatten_out(name, out, grid_stride=None):
    """store a reference return value (tensor or dict) under name/out/...; a large feature_grid is
    stored with a channel stride (key suffix '@c<stride>')"""
    keys = []
    if torch.is_tensor(out):
        put(name, "out", "ret", out); keys.append("ret")
        return keys
    for k, v in out.items():
        if k in ("xs", "ys", "covs", "sizes", "features") or v is None:
            continue
        if isinstance(v, dict):
            for kk, vv in v.items():
                put(name, "out", f"{k}/{kk}", vv); keys.append(f"{k}/{kk}")
        elif k == "feature_grid" and grid_stride:
            put(name, "out", f"{k}@c{grid_stride}", v[:, ::grid_stride]); keys.append(f"{k}@c{grid_stride}")
        else:
            put(name, "out", k, v); keys.append(k)
    return keys


def same(a, b):
    if torch.is_tensor(a):
        assert torch.equal(a, b), "aten_port differs from the reference"
    elif isinstance(a, dict):
        for k in a:
            if a[k] is not None and k not in ("xs", "ys", "covs", "sizes", "features"):
                same(a[k], b[k])


def add_render(name, blob, kwargs, subsample=None, note="", grid_stride=None):
    """one splat_features call; blob = dict of tensors; kwargs json-able (+ tensors for features/viz_colors)"""
    tens = {k: v for k, v in kwargs.items() if torch.is_tensor(v)}
    plain = {k: v for k, v in kwargs.items() if not torch.is_tensor(v) and not callable(v)}
    call = dict(kwargs)
    ref = REF.splat_features(**blob, **call)
    port_kw = {k: v for k, v in call.items()}
    port = aten_port.render(**blob, **port_kw)
    same(ref, port)
    for k, v in blob.items():
        put(name, "in", k, v)
    for k, v in tens.items():
        put(name, "in", k, v)
    if subsample:                       # big viz images: keep a strided view + global moments
        img = ref["feature_img"]
        put(name, "out", "feature_img_sub", img[..., ::subsample, ::subsample])
        put(name, "out", "feature_img_moments", torch.stack([img.sum(), (img * img).sum(), img.amax(), img.amin()]))
        outs = ["feature_img_sub", "feature_img_moments"]
    else:
        outs = flatten_out(name, ref, grid_stride)
    cases.append({"name": name, "func": "splat_features", "kwargs": plain, "tensor_kwargs": sorted(tens),
                  "uses_viz_score_fn": "viz_score_fn" in kwargs, "outs": outs, "subsample": subsample,
                  "note": note})


def tblob(d, dtype=None):
    out = {k: torch.from_numpy(np.asarray(v)) for k, v in d.items() if k != "features"}
    if dtype is not None:
        out = {k: (v.to(dtype) if k != "sizes" else v) for k, v in out.items()}
    return out


# ---- (i) the 40 real ellipses, script recipe, 64x64 d_score; fp64 as the scripts run it -------------
ellipses = []
for path in sorted(glob.glob("/root/reference/assets/results/demo/*/state/state.json")):
    demo = path.split("/")[-3]
    for i, item in enumerate(json.load(open(path))["ellipse_lists"]):
        ellipses.append({"demo": demo, "idx": i, "ellipse": item[0]})
assert len(ellipses) == 40, len(ellipses)
fg64 = []
for e in ellipses:
    (xc, yc), (d1, d2), ang = e["ellipse"]
    mean, cov = REF.ellipse_to_gaussian(xc, yc, d1 / 2, d2 / 2,
                                        np.radians((((180 - ang) % 180) + 90) % 180))
    ob = blob_oracle.blob_from_ellipse(e["ellipse"], 512, 512)
    nm, nc = mean / np.array([512, 512]), cov / (np.sqrt(512 ** 2 + 512 ** 2) ** 2)  # normalize_gs, blobctrl_inference.py:88-98
    assert np.array_equal(ob["covs"][0, 0], nc) and ob["xs"][0] == nm[0] and ob["ys"][0] == nm[1]
    blob = {"xs": torch.tensor(nm[0]).unsqueeze(0), "ys": torch.tensor(nm[1]).unsqueeze(0),
            "covs": torch.tensor(nc).unsqueeze(0).unsqueeze(0), "sizes": torch.tensor([1.0]).unsqueeze(0)}
    d = REF.splat_features(**blob, score_size=(64, 64), return_d_score=True)
    assert torch.equal(d, aten_port.render(**blob, score_size=(64, 64), return_d_score=True))
    assert d.shape == (1, 2, 64, 64) and d.dtype == torch.float64 and torch.isfinite(d).all()
    assert torch.equal(d[:, 0], 1 - d[:, 1])          # M=1: bg = 1 - fg exactly
    fg64.append(d[0, 1].numpy())
arrays["ellipses/fg64"] = np.stack(fg64)              # [40,64,64] fp64; bg = 1 - fg
json.dump(ellipses, open(os.path.join(HERE, "ellipses.json"), "w"), indent=0)

# ---- (ii) UI preview path at 512x512 (blobctrl_app.py:637-646), strided subsample --------------------
for idx in (1, 12, 30):                                # round, rotated, thin (replace_knife)
    ob = tblob(blob_oracle.blob_from_ellipse(ellipses[idx]["ellipse"], 512, 512))
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        b = {k: (v.to(dt) if k != "sizes" else v) for k, v in ob.items()}
        add_render(f"viz512_{idx}_{tag}", b,
                   dict(interp_size=64, viz_size=(512, 512), is_viz=True, ret_layout=True, score_size=64,
                        viz_score_fn=REF.viz_score_fn, viz_colors=REF.BLOB_VIS_COLORS, only_vis=True),
                   subsample=8, note=f"{ellipses[idx]['demo']}[{ellipses[idx]['idx']}]")
assert np.array_equal(REF.BLOB_VIS_COLORS.numpy(), blob_oracle.BLOB_VIS_COLORS)

# ---- (iii) seeded synthetic general [N,M] path -----------------------------------------------------
for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
    syn = blob_oracle.synthetic_blobs(2, 12, seed=1, c=24)
    b = tblob(syn, dt); f = torch.from_numpy(syn["features"]).to(dt)
    add_render(f"general_{tag}", b, dict(score_size=32, interp_size=8, features=f))
    add_render(f"general_noresize_{tag}", b, dict(score_size=16, interp_size=16, features=f, ret_layout=False))
    add_render(f"general_fg_{tag}", b, dict(score_size=16, return_d_score=True, only_splatting_fg=True))
    add_render(f"general_bg_{tag}", b, dict(score_size=16, return_d_score=True, only_splatting_bg=True))
    thin = blob_oracle.synthetic_blobs(2, 12, seed=2, thin=True, c=8)
    add_render(f"thin_{tag}", tblob(thin, dt), dict(score_size=32, return_d_score=True), note="numerics only")
    # generic [N,M] viz with per-batch colours and the composite-of-viz-scores branch
    add_render(f"general_viz_{tag}", b, dict(score_size=16, interp_size=16, viz_size=24, is_viz=True, only_vis=True,
                                              viz_score_fn=REF.viz_score_fn, viz_colors=REF.BLOB_VIS_COLORS))

# ---- (iv) edge cases ---------------------------------------------------------------------------------
syn = blob_oracle.synthetic_blobs(3, 5, seed=3, c=7)
b = tblob(syn, torch.float32); f = torch.from_numpy(syn["features"])
b["sizes"] = torch.tensor([[1, 0, 1, 0.49, 0.5], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1]], dtype=torch.float32)
add_render("edge_gate", b, dict(score_size=8, interp_size=8, features=f), note="sizes<0.5 -> 1e-6; all gated image")
b3 = dict(b); b3["sizes"] = b["sizes"].unsqueeze(-1)
add_render("edge_sizes_nm1", b3, dict(score_size=8, return_d_score=True), note="[N,M,1] sizes (utils.py:165-166)")
bo = dict(b); bo["xs"] = b["xs"] * 1.8 - 0.4; bo["ys"] = b["ys"] * 1.8 - 0.4
add_render("edge_outside", bo, dict(score_size=8, return_d_score=True), note="centres 40% outside the image")
deg = tblob(blob_oracle.blob_from_ellipse(ellipses[0]["ellipse"], 512, 512))
assert ellipses[0]["ellipse"][1] == [1e-05, 1e-05]
add_render("edge_degenerate_f64", deg, dict(score_size=(64, 64), return_d_score=True), note="det ~ 2e-33")
rect = tblob(blob_oracle.blob_from_ellipse(ellipses[30]["ellipse"], 512, 512))
add_render("edge_rect_f64", rect, dict(score_size=(48, 80), return_d_score=True), note="H != W tuple path")
add_render("edge_rect_f32", {k: (v.float()) for k, v in rect.items()}, dict(score_size=(48, 80), return_d_score=True))
syn1 = blob_oracle.synthetic_blobs(1, 1, seed=4, c=1024)
add_render("edge_pipeline_like", tblob(syn1, torch.float32),
           dict(score_size=64, interp_size=64, features=torch.from_numpy(syn1["features"]), ret_layout=False),
           note="K=2, C=1024 at latent res", grid_stride=128)

# ---- (v) stage 3 on its own incl. the bilinear resize branch (utils.py:70-73) -------------------------
g = torch.Generator().manual_seed(5)
for nm, (n, k, h, w, c, size, cl) in {
        "s3_plain": (2, 5, 12, 12, 9, 12, False), "s3_cl": (2, 5, 12, 12, 9, 12, True),
        "s3_down": (2, 5, 24, 24, 9, 16, False), "s3_up_cl": (1, 3, 16, 16, 4, 40, True),
        "s3_odd": (1, 4, 15, 15, 6, 7, False), "s3_none": (1, 4, 6, 10, 3, None, False),
        "s3_k1": (2, 1, 64, 64, 1024, 64, False)}.items():
    sc = torch.rand((n, h, w, k) if cl else (n, k, h, w), generator=g)
    ft = torch.randn((n, k, c), generator=g)
    ref = REF.splat_features_from_scores(sc, ft, size, channels_last=cl)
    assert torch.equal(ref, aten_port.feature_splat(sc, ft, size, cl))
    put(nm, "in", "scores", sc); put(nm, "in", "features", ft)
    if nm == "s3_k1":
        put(nm, "out", "ret@c128", ref[:, ::128]); outs = ["ret@c128"]
    else:
        put(nm, "out", "ret", ref); outs = ["ret"]
    cases.append({"name": nm, "func": "splat_features_from_scores", "kwargs": {"size": size, "channels_last": cl},
                  "outs": outs})

# ---- (vi) pyramid_resize (utils.py:280-294) -----------------------------------------------------------
for nm, (shape, cutoff) in {"pyr_64_8": ((2, 3, 64, 64), 8), "pyr_odd": ((1, 2, 20, 20), 3),
                            "pyr_noop": ((1, 2, 8, 8), 8)}.items():
    img = torch.rand(shape, generator=g)
    ref = REF.pyramid_resize(img, cutoff)
    port = aten_port.halve_pyramid(img, cutoff)
    assert ref.keys() == port.keys() and all(torch.equal(ref[k], port[k]) for k in ref)
    put(nm, "in", "img", img)
    for k, v in ref.items():
        put(nm, "out", str(k), v)
    cases.append({"name": nm, "func": "pyramid_resize", "kwargs": {"cutoff": cutoff}, "outs": [str(k) for k in ref]})

np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
json.dump({"torch": torch.__version__, "cases": cases}, open(os.path.join(HERE, "cases.json"), "w"), indent=1)
print(f"{len(cases)} cases, {len(arrays)} arrays, "
      f"{os.path.getsize(os.path.join(HERE, 'golden.npz')) / 1e6:.2f} MB")
