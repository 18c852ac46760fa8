"""GPU parity: the CUDA path (through the reference-signature API -> ctypes -> C ABI) against
(a) the golden fixtures recorded from the real reference and (b) the float64 numpy oracle on seeded
inputs, plus size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star; SURVEY.md §7.2 for the scale-relative definition) — never widened per test:
  fp32 maps : |got - ref| <= 1e-5 * max|ref|   [scores live in [0,1]]; thin blobs are compared with the float64
              reference on the same (float32-rounded) inputs, because the reference's own float32 run is 6.7e-6 abs
              away from its float64 self there
  bf16 maps : |got - ref| <= 1e-2 * max|ref|
  fp16 maps : 2e-3 (north_star names no f16 bar; 4 half-ulps of f16)
  fp64 maps : 1e-9 relative
  ordering / indexing (argmax_k, channel order, fg/bg selection): exact.
"""
import numpy as np
import pytest
import torch

import golden_util as G
from oracle import blob_oracle

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _impl():
    import blobctrl_b200.utils.utils as U
    return U


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _np(t):
    return t.detach().float().cpu().numpy() if t.dtype in (torch.bfloat16, torch.float16) else t.detach().cpu().numpy()


def _blob(syn, dtype=torch.float32):
    return {k: (_cuda(v).to(dtype) if k != "sizes" else _cuda(v)) for k, v in syn.items() if k != "features"}


def close_scaled(got, want, rel, what):
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: {got.shape} vs {want.shape}"
    scale = max(float(np.abs(want).max()), 1e-30)
    err = np.abs(got - want)
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    assert err.max() <= rel * scale, f"{what}: max abs err {err.max():.3e} > {rel:g} * {scale:.3g}"


CASES = G.cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_cases(case):
    U = _impl()
    f64 = not case["name"].endswith("_f32") and any(
        G.arrays()[f"{case['name']}/out/{k}"].dtype == np.float64 for k in case["outs"])
    for key, got, want in G.run_case(case, U, wrap=_cuda, to_np=_np, identity_fn=U.viz_score_fn):
        if key == "feature_img_moments":
            G.check_close(got, want, 1e-9 if f64 else 3e-5, 0, f"{case['name']}:{key}")
            continue
        assert got.dtype == want.dtype, (key, got.dtype, want.dtype)
        if case["name"] == "thin_f32":
            # thin blobs in fp32: the reference's own LU solve is 6.7e-6 abs / 5e-3 rel from its fp64 self, so the yardstick
            # is the float64 reference arithmetic on the same float32 inputs (oracle pinned to the reference at 1e-13)
            ins = G.case_inputs(case)
            want = blob_oracle.splat_features(**{k: ins[k] for k in ("xs", "ys", "covs", "sizes")}, score_size=32,
                                              return_d_score=True, dtype=np.float64)
        close_scaled(got, want, 1e-9 if f64 else 1e-5, f"{case['name']}:{key}")


def test_forty_demo_ellipses_fp64_script_recipe():
    """scripts/blobctrl_inference.py:71-117 on all 40 state.json ellipses, float64 like the scripts."""
    U = _impl()
    want = G.arrays()["ellipses/fg64"]
    for i, e in enumerate(G.ellipses()):
        mean, cov = U.get_gs_from_ellipse(e["ellipse"])
        nm, nc = U.normalize_gs(mean, cov, 512, 512)
        blob = U.get_blob_dict_from_norm_gs(nm, nc, device=DEV)
        score = U.get_blob_score_from_blob_dict(blob, score_size=(64, 64))
        assert score.shape == (2, 64, 64) and score.dtype == torch.float64
        got = score.cpu().numpy()
        close_scaled(got[1], want[i], 1e-9, f"{e['demo']}[{e['idx']}] fg")
        close_scaled(got[0], 1 - want[i], 1e-9, f"{e['demo']}[{e['idx']}] bg")
        if i < 8:
            # the scripts' own blob dict — host tensors, as blobctrl_inference.py:101-109 builds them — through the
            # import swap alone: uploaded, rendered on the GPU, same bits; the preview takes the host palette too
            host = U.get_blob_dict_from_norm_gs(nm, nc, device="cpu")
            hs = U.get_blob_score_from_blob_dict(host, score_size=(64, 64))
            assert hs.is_cuda and torch.equal(hs, score)
            hv = U.get_blob_vis_img_from_blob_dict(host, viz_size=(64, 64), score_size=(64, 64))
            assert hv.is_cuda and torch.equal(hv, U.get_blob_vis_img_from_blob_dict(blob, viz_size=(64, 64), score_size=(64, 64)))
        # same ellipse with float32 tensors: 1e-5 of scale against the float64 reference
        b32 = {k: (v.float()) for k, v in blob.items()}
        got32 = U.get_blob_score_from_blob_dict(b32, score_size=(64, 64)).cpu().numpy()
        assert got32.dtype == np.float32
        if e["ellipse"][1] != [1e-05, 1e-05]:
            close_scaled(got32[1], want[i], 1e-5, f"{e['demo']}[{e['idx']}] fg fp32")
        else:                                   # degenerate: exactly one pixel at 1.0, rest 0
            assert (got32[1] == 1.0).sum() == 1 and (got32[1] > 0).sum() == 1
            assert np.array_equal(got32[1] > 0, want[i] > 0)


@pytest.mark.parametrize("n,m,s,c,seed", [(1, 16, 64, 320, 1), (3, 32, 64, 320, 2), (2, 64, 64, 64, 3), (2, 5, 40, 33, 4)])
def test_synthetic_vs_fp64_oracle_fp32(n, m, s, c, seed):
    """BASELINE config 2 (16 blobs x 64x64 x 320ch) and neighbours, full dict, fp32 vs the fp64 oracle."""
    U = _impl()
    syn = blob_oracle.synthetic_blobs(n, m, seed=seed, c=c)
    want = blob_oracle.splat_features(**syn, score_size=s, interp_size=s, dtype=np.float64)
    for engine in ("fma", "auto"):
        got = U.splat_features(**_blob(syn), features=_cuda(syn["features"]), score_size=s, interp_size=s,
                               engine=engine)
        assert set(got) == set(want)
        assert got["feature_img"] is None and got["entropy_img"] is None
        close_scaled(_np(got["scores_pyramid"][s]), want["scores_pyramid"][s], 1e-5, "composed")
        close_scaled(_np(got["composed_scores"]), want["composed_scores"], 1e-5, "composed NHWK view")
        close_scaled(_np(got["raw_scores"]), want["raw_scores"], 1e-5, "raw")
        close_scaled(_np(got["feature_grid"]), want["feature_grid"], 1e-5, "feature grid")
        assert got["feature_grid"].is_contiguous() and got["feature_grid"].shape == (n, c, s, s)
        # ordering / indexing exact: front-most blob per pixel agrees wherever the oracle's margin is real
        d = want["scores_pyramid"][s]
        top2 = np.sort(d, axis=1)[:, -2:]
        sure = (top2[:, 1] - top2[:, 0]) > 1e-5
        assert np.array_equal(_np(got["scores_pyramid"][s]).argmax(1)[sure], d.argmax(1)[sure])
        no_layout = U.splat_features(**_blob(syn), features=_cuda(syn["features"]), score_size=s, interp_size=s,
                                     ret_layout=False, engine=engine)
        assert "raw_scores" not in no_layout
        close_scaled(_np(no_layout["feature_grid"]), want["feature_grid"], 1e-5, "feature grid (no layout)")
        close_scaled(_np(no_layout["scores_pyramid"][s]), want["scores_pyramid"][s], 1e-5, "composed (no layout)")


def test_thin_blobs_closer_to_fp64_than_reference_fp32():
    """Whitened quadratic form: on thin blobs the CUDA fp32 result is nearer the fp64 reference than the
    reference's own fp32 run (golden 'thin_f32' vs 'thin_f64')."""
    U = _impl()
    z = G.arrays()
    ins = {k: _cuda(z[f"thin_f32/in/{k}"]) for k in ("xs", "ys", "covs", "sizes")}
    got = _np(U.splat_features(**ins, score_size=32, return_d_score=True))
    ref64 = blob_oracle.splat_features(**{k: z[f"thin_f32/in/{k}"] for k in ("xs", "ys", "covs", "sizes")},
                                       score_size=32, return_d_score=True, dtype=np.float64)
    ours = np.abs(got - ref64).max()
    theirs = np.abs(z["thin_f32/out/ret"].astype(np.float64) - ref64).max()
    assert ours <= 1e-5 and ours <= theirs * 1.5 + 1e-7, (ours, theirs)


@pytest.mark.parametrize("dtype,rel", [(torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
def test_multiscale_config3_half(dtype, rel):
    """BASELINE config 3 at reduced batch: 32 blobs, levels 64/32/16/8, C = 320/640/1280/1280, 16-bit maps."""
    U = _impl()
    n, m = 2, 32
    syn = blob_oracle.synthetic_blobs(n, m, seed=1)
    rng = np.random.default_rng(11)
    chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
    feats = {s: rng.standard_normal((n, m + 1, c)).astype(np.float32) for s, c in chans.items()}
    d = blob_oracle.render_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], 64, 64, np.float64)
    pyr = blob_oracle.pyramid_resize(d, 8)
    got = U.splat_features_multiscale(**_blob(syn), score_size=64,
                                      level_features={s: _cuda(f).to(dtype) for s, f in feats.items()},
                                      out_dtype=dtype)
    assert sorted(got["scores_pyramid"]) == [8, 16, 32, 64]
    for s, c in chans.items():
        assert got["scores_pyramid"][s].dtype == dtype and got["feature_grids"][s].shape == (n, c, s, s)
        close_scaled(_np(got["scores_pyramid"][s]), pyr[s], rel, f"scores@{s}")
        f_q = _np(_cuda(feats[s]).to(dtype)).astype(np.float64)          # features as the kernel sees them
        want = blob_oracle.splat_features_from_scores(pyr[s], f_q, s, channels_last=False)
        close_scaled(_np(got["feature_grids"][s]), want, rel, f"grid@{s}")


def test_multiscale_fp32():
    U = _impl()
    syn = blob_oracle.synthetic_blobs(2, 32, seed=2)
    rng = np.random.default_rng(12)
    feats = {64: rng.standard_normal((2, 33, 64)).astype(np.float32), 16: rng.standard_normal((2, 33, 48)).astype(np.float32)}
    d = blob_oracle.render_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], 64, 64, np.float64)
    pyr = blob_oracle.pyramid_resize(d, 16)
    got = U.splat_features_multiscale(**_blob(syn), score_size=64, level_features={s: _cuda(f) for s, f in feats.items()})
    for s in (64, 32, 16):
        close_scaled(_np(got["scores_pyramid"][s]), pyr[s], 1e-5, f"scores@{s}")
    for s, f in feats.items():
        want = blob_oracle.splat_features_from_scores(pyr[s], f.astype(np.float64), s, channels_last=False)
        close_scaled(_np(got["feature_grids"][s]), want, 1e-5, f"grid@{s}")


@pytest.mark.parametrize("m", [1, 31, 32, 33, 64, 100])
def test_warp_scan_matches_lane_pixel(m):
    """The two stage-2 mappings agree to a few ulp (scan re-association), and both match the oracle."""
    U = _impl()
    syn = blob_oracle.synthetic_blobs(2, m, seed=5)
    b = _blob(syn)
    a = U.splat_features(**b, score_size=48, return_d_score=True, composite_mode="lane_pixel")
    w = U.splat_features(**b, score_size=48, return_d_score=True, composite_mode="warp_scan")
    want = blob_oracle.render_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], 48, 48, np.float64)
    close_scaled(_np(a), want, 1e-5, "lane_pixel")
    close_scaled(_np(w), want, 1e-5, "warp_scan")
    assert (a - w).abs().max().item() <= 2e-6
    for sel in ("only_splatting_fg", "only_splatting_bg"):
        ws = U.splat_features(**b, score_size=48, return_d_score=True, composite_mode="warp_scan", **{sel: True})
        ls = U.splat_features(**b, score_size=48, return_d_score=True, **{sel: True})
        assert ws.shape == ls.shape and (ws - ls).abs().max().item() <= 2e-6


def test_selection_and_layout_exact():
    """fg drops channel 0, bg keeps only channel 0 as [N,1,H,W]; channel k is blob k-1 — bit-exact slices."""
    U = _impl()
    syn = blob_oracle.synthetic_blobs(3, 7, seed=6)
    b = _blob(syn)
    full = U.splat_features(**b, score_size=20, return_d_score=True)
    fg = U.splat_features(**b, score_size=20, return_d_score=True, only_splatting_fg=True)
    bg = U.splat_features(**b, score_size=20, return_d_score=True, only_splatting_bg=True)
    assert full.shape == (3, 8, 20, 20) and fg.shape == (3, 7, 20, 20) and bg.shape == (3, 1, 20, 20)
    assert torch.equal(fg, full[:, 1:]) and torch.equal(bg, full[:, :1])
    # front-most = highest index: a lone opaque blob in the last slot hides everything beneath its centre
    assert torch.allclose(full.sum(1), torch.ones_like(full[:, 0]), atol=2e-6)   # partition of unity


def test_large_m_chunking_and_odd_sizes():
    """M > the 128-blob shared-memory chunk, W not a multiple of the vector width, tiny images."""
    U = _impl()
    for (n, m, h) in [(1, 300, 17), (2, 129, 6), (1, 2, 1)]:
        syn = blob_oracle.synthetic_blobs(n, m, seed=7)
        got = U.splat_features(**_blob(syn), score_size=h, return_d_score=True)
        want = blob_oracle.render_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], h, h, np.float64)
        close_scaled(_np(got), want, 1e-5, f"M={m} S={h}")


def test_full_size_properties_config5():
    """BASELINE config 5 at full size on one GPU (1024 images x 64 blobs, 64x64, C=320, fp32): properties
    that do not need the oracle — partition of unity, linearity of stage 3 in the features, constant
    features reproduce themselves — plus a 4-image slice against the fp64 oracle."""
    U = _impl()
    n, m, s, c = 1024, 64, 64, 320
    syn = blob_oracle.synthetic_blobs(n, m, seed=0, c=c)
    b = _blob(syn)
    f = _cuda(syn["features"])
    out = U.splat_features(**b, features=f, score_size=s, interp_size=s, ret_layout=False)
    d, g = out["scores_pyramid"][s], out["feature_grid"]
    assert d.shape == (n, m + 1, s, s) and g.shape == (n, c, s, s)
    assert (d.sum(1) - 1).abs().max().item() <= 4e-6                       # sum_k d_k = 1
    assert d.min().item() >= 0 and d.max().item() <= 1
    ones = U.splat_features_from_scores(d, torch.ones(n, m + 1, 8, device=DEV), s, channels_last=False)
    assert (ones - 1).abs().max().item() <= 4e-6                           # constant features -> constant map
    f2 = torch.randn_like(f)
    lin = U.splat_features_from_scores(d[:64], (2 * f + 3 * f2)[:64], s, channels_last=False)
    sep = 2 * U.splat_features_from_scores(d[:64], f[:64], s, channels_last=False) + \
        3 * U.splat_features_from_scores(d[:64], f2[:64], s, channels_last=False)
    assert (lin - sep).abs().max().item() <= 1e-4 * float(lin.abs().max())   # linearity
    sl = slice(509, 513)
    want = blob_oracle.splat_features(**{k: v[sl] for k, v in syn.items()}, score_size=s, interp_size=s,
                                      dtype=np.float64, ret_layout=False)
    close_scaled(_np(d[sl]), want["scores_pyramid"][s], 1e-5, "cfg5 composed slice")
    close_scaled(_np(g[sl]), want["feature_grid"], 1e-5, "cfg5 grid slice")


def _have_reference():
    import os
    return os.path.isdir(os.path.join(os.path.dirname(G.HERE), "baseline", "_ref", "blobctrl"))


@pytest.mark.skipif(not _have_reference(), reason="baseline/_ref not installed (scripts/install_reference.sh)")
def test_full_size_config5_whole_batch_against_the_installed_reference():
    """BASELINE config 5 at full size, EVERY image: the fused float32 render against the UNMODIFIED reference
    splat_features (baseline/_ref, blobctrl/utils/utils.py:80-241) run on this GPU in float64 on the same float32
    inputs, 256 images at a time — 1e-5 of scale on the composed maps and on the feature grid."""
    from baseline import ref_loader
    R = ref_loader.load(need_pipeline=False).utils
    U = _impl()
    n, m, s, c = 1024, 64, 64, 320
    syn = blob_oracle.synthetic_blobs(n, m, seed=0, c=c)
    b = _blob(syn)
    f = _cuda(syn["features"])
    out = U.splat_features(**b, features=f, score_size=s, interp_size=s, ret_layout=False)
    d, g = out["scores_pyramid"][s], out["feature_grid"]
    worst_d = worst_g = 0.0
    for lo in range(0, n, 256):
        sl = slice(lo, lo + 256)
        ref = R.splat_features(**{k: v[sl].double() for k, v in b.items()}, features=f[sl].double(), score_size=s,
                               interp_size=s, ret_layout=False)
        rd, rg = ref["scores_pyramid"][s], ref["feature_grid"]
        assert rd.dtype == torch.float64 and rd.shape == d[sl].shape and rg.shape == g[sl].shape
        worst_d = max(worst_d, (d[sl].double() - rd).abs().max().item())                       # scores live in [0, 1]
        worst_g = max(worst_g, ((g[sl].double() - rg).abs().amax(dim=(1, 2, 3)) / rg.abs().amax(dim=(1, 2, 3))).max().item())
        am_ours, am_ref = d[sl].argmax(1), rd.argmax(1)                                          # front-most visible blob per pixel
        top2 = rd.topk(2, dim=1).values
        sure = (top2[:, 0] - top2[:, 1]) > 1e-5
        assert torch.equal(am_ours[sure], am_ref[sure])
        del ref, rd, rg
    assert worst_d <= 1e-5 and worst_g <= 1e-5, (worst_d, worst_g)


def test_pipeline_conditioning_entry():
    """pipeline_blobnet.py:973-984 + :724-739: K=1, C=1024, fp16, layouts [2B,1029,64,128] / [2B,5,64,128]."""
    from blobctrl_b200.pipelines import construct_blobnet_input, prepare_blob_conditioning
    U = _impl()
    e = G.ellipses()[20]["ellipse"]
    nm, nc = U.normalize_gs(*U.get_gs_from_ellipse(e), 512, 512)
    gs = U.get_blob_score_from_blob_dict(U.get_blob_dict_from_norm_gs(nm, nc, device="cuda"), (64, 64)).unsqueeze(0)
    g = torch.Generator(device="cpu").manual_seed(3)
    dino = torch.randn(1, 1, 1024, generator=g).to(DEV)
    cond = prepare_blob_conditioning(gs, dino, batch=4, dtype=torch.float16, device=DEV)
    assert cond.fg_gs_feats.shape == (4, 1024, 64, 64) and cond.fg_gs_feats.dtype == torch.float16
    want = (gs[:, 1:2].half().float()[:, :, None] * dino.half().float()[:, 0, :, None, None, None]).squeeze(3)
    want = want.reshape(1, 1024, 64, 64).expand(4, -1, -1, -1)
    assert (cond.fg_gs_feats.float() - want).abs().max().item() <= 1e-2 * float(want.abs().max())
    lat = torch.randn(4, 4, 64, 64, device=DEV, dtype=torch.float16)
    img = torch.randn(4, 4, 64, 64, device=DEV, dtype=torch.float16)
    x = construct_blobnet_input(lat, cond.fg_gs_scores, img, cond.fg_gs_feats)
    assert x.shape == (4, 1029, 64, 128)
    assert torch.equal(x[:, :4, :, :64], img) and torch.equal(x[:, :4, :, 64:], lat)
    assert torch.equal(x[:, 4:5, :, :64], cond.fg_gs_scores) and torch.equal(x[:, 5:, :, 64:], cond.fg_gs_feats)
    xb = construct_blobnet_input(lat, cond.bg_gs_scores, img, background=True)
    assert xb.shape == (4, 5, 64, 128)


def test_errors_and_no_cpu_fallback():
    U = _impl()
    syn = blob_oracle.synthetic_blobs(2, 3, seed=9, c=4)
    cpu = {k: torch.from_numpy(v) for k, v in syn.items() if k != "features"}
    with pytest.raises(RuntimeError, match="no CPU"):       # the stage functions take device maps only
        U.splat_features_from_scores(torch.rand(2, 4, 8, 8), torch.from_numpy(syn["features"]), 8, channels_last=False)
    with pytest.raises(RuntimeError, match="no CPU"):
        U.pyramid_resize(torch.rand(2, 4, 8, 8), 4)
    b = _blob(syn)
    # the renderer takes the scripts' host-built blob dict: parameters are uploaded, maps are rendered and stay on the GPU
    up = U.splat_features(**cpu, score_size=8, return_d_score=True)
    assert up.is_cuda and torch.equal(up, U.splat_features(**b, score_size=8, return_d_score=True))
    with pytest.raises(RuntimeError):                       # tuple size with N*M > 1 (utils.py:157-159)
        U.splat_features(**b, score_size=(8, 8), return_d_score=True)
    with pytest.raises(TypeError):                          # interp_size missing (utils.py:291)
        U.splat_features(**b, score_size=8, features=_cuda(syn["features"]))
    with pytest.raises(KeyError):                           # interp_size not a pyramid level
        U.splat_features(**b, score_size=8, interp_size=16, features=_cuda(syn["features"]))
    with pytest.raises(RuntimeError):                       # feature rows != channels
        U.splat_features(**b, score_size=8, interp_size=8, features=_cuda(syn["features"])[:, :2])
    sing = {k: v.clone() for k, v in b.items()}
    sing["covs"][1, 2] = torch.tensor([[0.01, 0.02], [0.02, 0.04]], device=DEV)              # rank 1
    with pytest.raises(RuntimeError, match="singular"):      # the reference's torch.linalg.solve raises (utils.py:143); opt-in here
        U.splat_features(**sing, score_size=8, return_d_score=True, check_singular=True)
    assert U.splat_features(**b, score_size=8, return_d_score=True, check_singular=True).shape == (2, 4, 8, 8)
    from blobctrl_b200 import _capi
    with pytest.raises(_capi.BlobSplatError):               # warp-scan limit
        big = blob_oracle.synthetic_blobs(1, 300, seed=1)
        U.splat_features(**_blob(big), score_size=8, return_d_score=True, composite_mode="warp_scan")


def test_cuda_graph_capture_and_streams():
    """The C ABI neither allocates nor synchronises: a render can be captured and replayed."""
    from blobctrl_b200 import ops
    syn = blob_oracle.synthetic_blobs(2, 16, seed=10, c=32)
    b = _blob(syn); f = _cuda(syn["features"])
    xs, ys, covs, sizes, n, m = ops.canonical_blobs(**b)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ref_d, _ = ops.render_scores(xs, ys, covs, sizes, 32, 32)
        ref_g = ops.feature_splat(ref_d, f)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        d, _ = ops.render_scores(xs, ys, covs, sizes, 32, 32)
        g = ops.feature_splat(d, f)
    d.zero_(); g.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(d, ref_d) and torch.equal(g, ref_g)


def test_custom_viz_score_fn_branch():
    """utils.py:199-209: a non-identity viz_score_fn re-composites the modified raw scores."""
    U = _impl()
    syn = blob_oracle.synthetic_blobs(2, 6, seed=12)
    fn_t = lambda s: torch.cat([s[..., :1], (s[..., 1:] * 1.5).clamp(max=1)], -1)
    fn_n = lambda s: np.concatenate([s[..., :1], np.minimum(s[..., 1:] * 1.5, 1)], -1)
    got = U.splat_features(**_blob(syn), score_size=16, interp_size=16, viz_size=16, is_viz=True, only_vis=True,
                           viz_score_fn=fn_t, viz_colors=U.BLOB_VIS_COLORS)
    want = blob_oracle.splat_features(**syn, score_size=16, interp_size=16, viz_size=16, is_viz=True, only_vis=True,
                                      viz_score_fn=fn_n, viz_colors=blob_oracle.BLOB_VIS_COLORS, dtype=np.float64)
    close_scaled(_np(got["feature_img"]), want["feature_img"], 1e-5, "custom viz")


@pytest.mark.parametrize("n,k,s,c,dtype,rel", [
    (3, 33, 64, 320, torch.float32, 1e-5), (2, 33, 32, 640, torch.float32, 1e-5), (2, 33, 8, 1280, torch.float32, 1e-5),
    (2, 65, 16, 1280, torch.bfloat16, 1e-2), (4, 17, 24, 96, torch.float16, 2e-3), (1, 128, 20, 64, torch.float32, 1e-5),
    (2, 12, 64, 64, torch.float32, 1e-5)])
def test_stage3_tensor_engine_vs_oracle(n, k, s, c, dtype, rel):
    """splat_features_from_scores on the tcgen05 engine (weights loaded from score maps into TMEM), both layouts,
    against the float64 oracle and against the FMA engine."""
    U = _impl()
    g = torch.Generator().manual_seed(n * 1000 + k)
    sc = torch.rand(n, k, s, s, generator=g)
    sc = sc / sc.sum(1, keepdim=True)                                   # convex weights like composed scores
    ft = torch.randn(n, k, c, generator=g)
    sc_d, ft_d = sc.to(DEV).to(dtype), ft.to(DEV).to(dtype)
    want = blob_oracle.splat_features_from_scores(_np(sc_d).astype(np.float64), _np(ft_d).astype(np.float64), s,
                                                  channels_last=False)
    got_t = U.splat_features_from_scores(sc_d, ft_d, s, channels_last=False, engine="tensor")
    got_f = U.splat_features_from_scores(sc_d, ft_d, s, channels_last=False, engine="fma")
    assert got_t.shape == (n, c, s, s) and got_t.dtype == dtype and got_t.is_contiguous()
    close_scaled(_np(got_t), want, rel, "tensor engine")
    close_scaled(_np(got_f), want, rel, "fma engine")
    nhwk = sc_d.permute(0, 2, 3, 1).contiguous()                        # true channels-last storage (strided planes)
    got_cl = U.splat_features_from_scores(nhwk, ft_d, s, channels_last=True, engine="tensor")
    close_scaled(_np(got_cl), want, rel, "tensor engine, NHWK")
    auto = U.splat_features_from_scores(sc_d, ft_d, s, channels_last=False)
    close_scaled(_np(auto), want, rel, "auto engine")


def test_tensor_engine_rejects_outside_envelope():
    U = _impl()
    from blobctrl_b200 import _capi
    sc = torch.rand(1, 5, 8, 8, device=DEV); ft = torch.randn(1, 5, 30, device=DEV)
    with pytest.raises(_capi.BlobSplatError):                           # float64 stays on the FMA engine
        U.splat_features_from_scores(sc.double(), ft.double(), 8, channels_last=False, engine="tensor")
    big = torch.rand(1, 130, 8, 8, device=DEV)
    with pytest.raises(_capi.BlobSplatError):                           # more than 128 planes
        U.splat_features_from_scores(big, torch.randn(1, 130, 64, device=DEV), 8, channels_last=False, engine="tensor")
    forced = U.splat_features_from_scores(sc, ft, 8, channels_last=False, engine="tensor")   # any C: ragged channel tile
    out = U.splat_features_from_scores(sc, ft, 8, channels_last=False)  # auto -> FMA (tiny K, C)
    assert (forced - out).abs().max().item() <= 1e-5 * out.abs().max().item()
    want = torch.einsum("nkhw,nkc->nchw", sc, ft)
    assert (out - want).abs().max().item() <= 1e-5 * want.abs().max().item()


def test_host_renderer_chunked_copy_matches_single_call():
    """blobctrl_b200.streaming: chunked H2D overlapped with the render == one splat_features call (images are
    independent), repeated calls reuse the buffers safely."""
    from blobctrl_b200.streaming import HostRenderer
    U = _impl()
    n, m, s, c = 37, 16, 32, 64
    r = HostRenderer(n, m, s, c, torch.float32, DEV, chunks=5)
    for seed in (21, 22):
        syn = blob_oracle.synthetic_blobs(n, m, seed=seed, c=c)
        host = {k: torch.from_numpy(v).pin_memory() for k, v in syn.items()}
        out = r(host["xs"], host["ys"], host["covs"], host["sizes"], host["features"])
        ref = U.splat_features(**_blob(syn), features=_cuda(syn["features"]), score_size=s, interp_size=s, ret_layout=False)
        torch.cuda.synchronize()
        assert torch.equal(out["scores_pyramid"][s], ref["scores_pyramid"][s])
        assert torch.equal(out["feature_grid"], ref["feature_grid"])


def test_host_renderer_back_to_back_calls_do_not_mix_inputs():
    """Double-buffered input staging: consecutive calls with different host inputs (copies of call i+1 overlap the renders
    of call i) each return their own result; outputs are snapshotted in stream order before the next call overwrites them."""
    from blobctrl_b200.streaming import HostRenderer
    U = _impl()
    n, m, s, c = 24, 9, 32, 64
    r = HostRenderer(n, m, s, c, torch.float32, DEV, chunks=3)
    sets = [blob_oracle.synthetic_blobs(n, m, seed=40 + i, c=c) for i in range(5)]
    hosts = [{k: torch.from_numpy(v).pin_memory() for k, v in syn.items()} for syn in sets]
    snaps = []
    for h in hosts:                                     # no synchronisation between the calls
        out = r(h["xs"], h["ys"], h["covs"], h["sizes"], h["features"])
        snaps.append((out["scores_pyramid"][s].clone(), out["feature_grid"].clone()))
    torch.cuda.synchronize()
    for syn, (d, g) in zip(sets, snaps):
        ref = U.splat_features(**_blob(syn), features=_cuda(syn["features"]), score_size=s, interp_size=s, ret_layout=False)
        assert torch.equal(d, ref["scores_pyramid"][s]) and torch.equal(g, ref["feature_grid"])


def test_ellipse_front_end_matches_script_recipe():
    """blobsplat_scores_ellipse (N3): all 40 demo ellipses as one batch == the reference's host recipe + renderer
    (golden fp64 maps), including the two degenerate 1e-5-pixel ellipses (exactly one pixel at 1.0)."""
    U = _impl()
    ell = G.ellipses()
    want = G.arrays()["ellipses/fg64"]
    batch = torch.tensor([[U.flatten_cv_ellipse(e["ellipse"])] for e in ell], dtype=torch.float32)   # [40, 1, 5]
    d = U.splat_ellipses(batch.to(DEV), image_size=(512, 512), score_size=64)
    assert d.shape == (40, 2, 64, 64) and d.dtype == torch.float32
    got = _np(d)
    for i, e in enumerate(ell):
        if e["ellipse"][1] == [1e-05, 1e-05]:
            assert (got[i, 1] == 1.0).sum() == 1 and (got[i, 1] > 0).sum() == 1 and np.array_equal(got[i, 1] > 0, want[i] > 0)
        else:
            close_scaled(got[i, 1], want[i], 1e-5, f"{e['demo']}[{e['idx']}] fg")
            close_scaled(got[i, 0], 1 - want[i], 1e-5, f"{e['demo']}[{e['idx']}] bg")


def test_ellipse_front_end_multi_blob_and_rect():
    """Several ellipses per image (depth order = index), existence flags, H != W render size, non-square source image."""
    U = _impl()
    rng = np.random.default_rng(5)
    n, m, img_h, img_w, h, w = 3, 6, 384, 640, 40, 72
    ell = np.stack([rng.uniform(0, img_w, (n, m)), rng.uniform(0, img_h, (n, m)), rng.uniform(20, 300, (n, m)),
                    rng.uniform(20, 300, (n, m)), rng.uniform(0, 180, (n, m))], -1)
    ell = ell.astype(np.float32).astype(np.float64)      # the entry point takes float32 ellipses: same inputs on both sides
    sizes = (rng.random((n, m)) > 0.2).astype(np.float32)
    xs = np.zeros((n, m)); ys = np.zeros((n, m)); covs = np.zeros((n, m, 2, 2))
    for i in range(n):
        for j in range(m):
            e = ((ell[i, j, 0], ell[i, j, 1]), (ell[i, j, 2], ell[i, j, 3]), ell[i, j, 4])
            mean, cov = blob_oracle.gs_from_ellipse(e)
            nm, nc = blob_oracle.normalize_gs(mean, cov, img_w, img_h)
            xs[i, j], ys[i, j], covs[i, j] = nm[0], nm[1], nc
    raw = blob_oracle.raw_scores(xs, ys, covs, sizes, h, w, np.float64)
    _, dref = blob_oracle.composite(raw)
    want = np.moveaxis(dref, -1, 1)
    got = U.splat_ellipses(torch.from_numpy(ell).float().to(DEV), torch.from_numpy(sizes).to(DEV), image_size=(img_h, img_w),
                           score_size=(h, w))
    assert got.shape == (n, m + 1, h, w)
    close_scaled(_np(got), want, 1e-5, "multi-blob ellipses")
    fg = U.splat_ellipses(torch.from_numpy(ell).float().to(DEV), torch.from_numpy(sizes).to(DEV), image_size=(img_h, img_w),
                          score_size=(h, w), only_splatting_fg=True)
    assert torch.equal(fg, got[:, 1:])


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
def test_fused_construct_blobnet_input_bit_exact(dtype):
    """N2: persistent canvases filled by blobsplat_conditioning_fill == the reference's per-step torch.cat layout
    (pipeline_blobnet.py:724-739) with fg_gs_feats from the stage-3 splat (:984), bit for bit."""
    from blobctrl_b200.pipelines import BlobNetInputBuffers, construct_blobnet_input, prepare_blob_conditioning
    U = _impl()
    b2, h, w, c = 4, 64, 64, 1024
    e = G.ellipses()[12]["ellipse"]
    gs = U.splat_ellipses([[U.flatten_cv_ellipse(e)]], image_size=(512, 512), score_size=64)          # [1,2,64,64]
    g = torch.Generator().manual_seed(7)
    dino = torch.randn(1, 1, c, generator=g).to(DEV)
    cond = prepare_blob_conditioning(gs, dino, batch=b2, dtype=dtype, device=DEV)
    fg_lat = torch.randn(b2, 4, h, w, generator=g).to(DEV).to(dtype)
    bg_lat = torch.randn(b2, 4, h, w, generator=g).to(DEV).to(dtype)
    bufs = BlobNetInputBuffers(b2, h, w, c, dtype, DEV)
    bufs.fill_static(cond.fg_gs_scores, cond.bg_gs_scores, dino.repeat(b2, 1, 1), fg_lat, bg_lat)
    for step in range(2):
        lat = torch.randn(b2, 4, h, w, generator=g).to(DEV).to(dtype)
        x, xb = bufs.update(lat)
        want = construct_blobnet_input(lat, cond.fg_gs_scores, fg_lat, cond.fg_gs_feats)
        want_bg = construct_blobnet_input(lat, cond.bg_gs_scores, bg_lat, background=True)
        assert x.shape == (b2, 4 + 1 + c, h, 2 * w) and xb.shape == (b2, 5, h, 2 * w)
        assert torch.equal(x, want) and torch.equal(xb, want_bg)


def test_graphed_renderers_match_eager():
    """blobctrl_b200.preview: CUDA-graph replay == the eager API (UI preview path and a cfg2-shaped render)."""
    from blobctrl_b200.preview import GraphedBlobRenderer, preview_renderer
    U = _impl()
    r = preview_renderer((128, 128), DEV)
    for idx in (1, 12, 30):
        blob = blob_oracle.blob_from_ellipse(G.ellipses()[idx]["ellipse"], 512, 512)
        b32 = {k: _cuda(v).float() for k, v in blob.items()}
        scores, img = r(blob["xs"], blob["ys"], blob["covs"])          # host float64 numpy in, like the app
        want = U.splat_features(**b32, interp_size=64, viz_size=(128, 128), is_viz=True, score_size=64,
                                viz_score_fn=U.viz_score_fn, viz_colors=U.BLOB_VIS_COLORS, only_vis=True)["feature_img"]
        torch.cuda.synchronize()
        assert img.shape == (1, 3, 128, 128) and (img - want).abs().max().item() <= 2e-6
    syn = blob_oracle.synthetic_blobs(1, 16, seed=3, c=320)
    g = GraphedBlobRenderer(1, 16, (64, 64), channels=320, device=DEV)
    for seed in (3, 4):
        syn = blob_oracle.synthetic_blobs(1, 16, seed=seed, c=320)
        sc, grid = g(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], _cuda(syn["features"]))
        ref = U.splat_features(**_blob(syn), features=_cuda(syn["features"]), score_size=64, interp_size=64, ret_layout=False)
        torch.cuda.synchronize()
        assert torch.equal(sc, ref["scores_pyramid"][64]) and torch.equal(grid, ref["feature_grid"])


def test_empty_and_degenerate_shapes():
    """No blobs (M = 0): the background owns every pixel; no images (N = 0): empty outputs, no launch."""
    U = _impl()
    xs = torch.zeros(2, 0, device=DEV); covs = torch.zeros(2, 0, 2, 2, device=DEV); sizes = torch.zeros(2, 0, device=DEV)
    d = U.splat_features(xs, xs, covs, sizes, score_size=8, return_d_score=True)
    assert d.shape == (2, 1, 8, 8) and torch.equal(d, torch.ones_like(d))
    e = U.splat_features(torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, 2, 2, device=DEV),
                         torch.zeros(0, 3, device=DEV), score_size=8, return_d_score=True)
    assert e.shape == (0, 4, 8, 8)
    g = U.splat_features_from_scores(torch.zeros(0, 4, 8, 8, device=DEV), torch.zeros(0, 4, 5, device=DEV), 8, channels_last=False)
    assert g.shape == (0, 5, 8, 8)


def test_randomised_shapes_vs_oracle():
    """Seeded sweep over ragged shapes (any N, M, H != W, C) through the general [N,M] renderer and both stage-3 engines."""
    from blobctrl_b200 import ops
    rng = np.random.default_rng(2024)
    for it in range(24):
        n, m = int(rng.integers(1, 4)), int(rng.integers(1, 41))
        h, w = int(rng.integers(1, 41)), int(rng.integers(1, 41))
        c = int(rng.choice([1, 3, 7, 32, 64, 96]))
        syn = blob_oracle.synthetic_blobs(n, m, seed=100 + it, thin=bool(it % 5 == 0), c=c)
        raw = blob_oracle.raw_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], h, w, np.float64)
        _, dref = blob_oracle.composite(raw)
        want_d = np.moveaxis(dref, -1, 1)
        b = _blob(syn)
        d, r = ops.render_scores(b["xs"], b["ys"], b["covs"], b["sizes"], h, w, want_raw=True)
        close_scaled(_np(d), want_d, 1e-5, f"case {it} composed {n}x{m}x{h}x{w}")
        close_scaled(_np(r)[:, 1:], np.moveaxis(raw, -1, 1), 1e-5, f"case {it} raw")
        want_g = blob_oracle.splat_features_from_scores(want_d, syn["features"].astype(np.float64), None, channels_last=False)
        for eng in ("fma", "auto"):
            g = ops.feature_splat(d, _cuda(syn["features"]), engine=eng)
            close_scaled(_np(g), want_g, 1e-5, f"case {it} grid C={c} {eng}")


def test_input_forms_noncontiguous_half_params_and_large_image():
    U = _impl()
    syn = blob_oracle.synthetic_blobs(2, 5, seed=31)
    want = blob_oracle.render_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], 16, 16, np.float64)
    covs_nc = _cuda(syn["covs"]).permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)       # non-contiguous view
    assert not covs_nc.is_contiguous()
    d = U.splat_features(_cuda(syn["xs"]), _cuda(syn["ys"]), covs_nc, _cuda(syn["sizes"]), score_size=16, return_d_score=True)
    close_scaled(_np(d), want, 1e-5, "non-contiguous covs")
    # bfloat16 parameters: the reference cannot run them at all ("lu_cpu" not implemented); here they are upcast
    h = {k: (_cuda(v).to(torch.bfloat16) if k != "sizes" else _cuda(v)) for k, v in syn.items()}
    dh = U.splat_features(**h, score_size=16, return_d_score=True)
    assert dh.dtype == torch.bfloat16
    want_h = blob_oracle.render_scores(_np(h["xs"]), _np(h["ys"]), _np(h["covs"]), syn["sizes"], 16, 16, np.float64)
    close_scaled(_np(dh), want_h, 1e-2, "bf16 parameters")
    # one blob on a 1024 x 1024 canvas, float32 and float64 (tuple-size path)
    one = blob_oracle.blob_from_ellipse(((500.0, 300.0), (400.0, 150.0), 33.0), 1024, 1024)
    w64 = blob_oracle.splat_features(**one, score_size=(1024, 1024), return_d_score=True)
    g64 = U.splat_features(**{k: _cuda(v) for k, v in one.items()}, score_size=(1024, 1024), return_d_score=True)
    assert g64.dtype == torch.float64
    close_scaled(_np(g64), w64, 1e-9, "1024^2 fp64")
    g32 = U.splat_features(**{k: _cuda(v).float() for k, v in one.items()}, score_size=(1024, 1024), return_d_score=True)
    close_scaled(_np(g32), w64, 1e-5, "1024^2 fp32")


def test_full_size_config3_properties():
    """BASELINE config 3 at full batch (N=64, 32 blobs, levels 64/32/16/8, C=320/640/1280/1280, bf16): shapes, dtype,
    partition of unity at every level, and agreement of a 2-image slice with the float64 oracle."""
    U = _impl()
    n, m = 64, 32
    syn = blob_oracle.synthetic_blobs(n, m, seed=0)
    chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
    g = torch.Generator().manual_seed(1)
    feats = {s: torch.randn(n, m + 1, c, generator=g).to(DEV).to(torch.bfloat16) for s, c in chans.items()}
    out = U.splat_features_multiscale(**_blob(syn), score_size=64, level_features=feats, out_dtype=torch.bfloat16)
    for s, c in chans.items():
        sc, gr = out["scores_pyramid"][s], out["feature_grids"][s]
        assert sc.shape == (n, m + 1, s, s) and gr.shape == (n, c, s, s) and gr.dtype == torch.bfloat16
        assert (sc.float().sum(1) - 1).abs().max().item() <= 2e-2
    sl = slice(30, 32)
    d = blob_oracle.render_scores(syn["xs"][sl], syn["ys"][sl], syn["covs"][sl], syn["sizes"][sl], 64, 64, np.float64)
    pyr = blob_oracle.pyramid_resize(d, 8)
    for s in chans:
        close_scaled(_np(out["scores_pyramid"][s][sl]), pyr[s], 1e-2, f"cfg3 scores@{s}")
        want = blob_oracle.splat_features_from_scores(pyr[s], _np(feats[s][sl]).astype(np.float64), s, channels_last=False)
        close_scaled(_np(out["feature_grids"][s][sl]), want, 1e-2, f"cfg3 grid@{s}")
    # the WHOLE batch (every CTA's item range of the TMA engine, every tile pair of the fused pyramid): each grid against the
    # float64 contraction of the produced 16-bit maps, each pyramid level against the pyramid kernel on the level above
    for s in chans:
        want = torch.einsum("nkhw,nkc->nchw", out["scores_pyramid"][s].double(), feats[s].double())
        err = (out["feature_grids"][s].double() - want).abs().max().item()
        assert err <= 1e-2 * want.abs().max().item(), f"cfg3 full batch grid@{s}: {err}"
    from blobctrl_b200 import ops
    pyr_k = ops.halving_pyramid(out["scores_pyramid"][64], 8)
    for s in (32, 16, 8):
        assert torch.equal(out["scores_pyramid"][s], pyr_k[s]), f"cfg3 full batch pyramid@{s}"


@pytest.mark.skipif(not _have_reference(), reason="baseline/_ref not installed (scripts/install_reference.sh)")
def test_full_size_config3_whole_batch_against_the_installed_reference():
    """BASELINE config 3, every image and level: bf16 maps of the one-call multi-scale render against the UNMODIFIED
    reference on this GPU in float64 — splat_features for the 64x64 maps, its pyramid_resize for the lower levels, its
    splat_features_from_scores per level (utils.py:80-241, 280-294, 57-77) — at 1e-2 of scale (the reference cannot run bf16)."""
    from baseline import ref_loader
    R = ref_loader.load(need_pipeline=False).utils
    U = _impl()
    n, m = 64, 32
    syn = blob_oracle.synthetic_blobs(n, m, seed=0)
    chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
    g = torch.Generator().manual_seed(1)
    feats = {s: torch.randn(n, m + 1, c, generator=g).to(DEV).to(torch.bfloat16) for s, c in chans.items()}
    b = _blob(syn)
    out = U.splat_features_multiscale(**b, score_size=64, level_features=feats, out_dtype=torch.bfloat16)
    rd = R.splat_features(**{k: v.double() for k, v in b.items()}, score_size=64, return_d_score=True)
    pyr = R.pyramid_resize(rd, cutoff=8)
    assert sorted(pyr) == [8, 16, 32, 64]
    for s in chans:
        err = (out["scores_pyramid"][s].double() - pyr[s]).abs().max().item()
        assert err <= 1e-2, f"cfg3 scores@{s} vs reference: {err}"
        want = R.splat_features_from_scores(pyr[s], feats[s].double(), s, channels_last=False)
        err = (out["feature_grids"][s].double() - want).abs().max().item()
        assert err <= 1e-2 * want.abs().max().item(), f"cfg3 grid@{s} vs reference: {err}"


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_residual_injection_bit_exact(dtype):
    """N4: fused scale + right-half slice + add == the reference's three elementwise passes (blobnet.py:936-938,
    pipeline_blobnet.py:1085-1087, unet_2d_condition.py:1215-1219), bit for bit, for float and per-sample scales."""
    from blobctrl_b200.pipelines import inject_residual
    g = torch.Generator().manual_seed(11)
    b, c, h = 4, 37, 16
    hidden = torch.randn(b, c, h, 2 * h, generator=g).to(DEV).to(dtype)
    res = torch.randn(b, c, h, 2 * h, generator=g).to(DEV).to(dtype)
    for scale in (1.2, torch.tensor([0.5, 1.0, 1.2, 0.0]).to(DEV).to(dtype)):
        s = scale if not torch.is_tensor(scale) else scale[:, None, None, None]
        scaled = res * s                                             # blobnet.py:936-938
        cropped = scaled[..., -scaled.shape[-2]:]                    # pipeline_blobnet.py:1085-1087
        want = hidden.clone()
        want[..., -want.shape[-2]:] = want[..., -want.shape[-2]:] + cropped     # unet_2d_condition.py:1218
        got = inject_residual(hidden.clone(), res, scale)
        assert torch.equal(got, want)
    sq = torch.randn(2, 8, h, h, generator=g).to(DEV).to(dtype); rq = torch.randn(2, 8, h, h, generator=g).to(DEV).to(dtype)
    assert torch.equal(inject_residual(sq.clone(), rq, 0.7), sq + rq * 0.7)    # square map: whole width (:1216)


@pytest.mark.parametrize("h,w,c,dtype,rel", [
    (7, 9, 64, torch.float32, 1e-5),       # odd pixel count: no 2-pixel stores, partial tile
    (6, 11, 96, torch.float32, 1e-5),      # even pixel count, not a multiple of 4, partial tile
    (20, 20, 352, torch.float32, 1e-5),    # two channel chunks, ragged second chunk (generic drain)
    (16, 16, 320, torch.float32, 1e-5),    # plane-stride specialisation 256
    (7, 9, 64, torch.bfloat16, 1e-2), (10, 13, 96, torch.float16, 2e-3)])
def test_fused_render_epilogue_paths(h, w, c, dtype, rel):
    """blobsplat_render's drain paths: pixel-pair vector stores (float maps), the per-lane 16-bit path and the generic
    path, on ragged tiles, plus an output buffer that is only 4-byte aligned (vector stores must not be used)."""
    from blobctrl_b200 import ops
    n, m = 3, 19
    syn = blob_oracle.synthetic_blobs(n, m, seed=h * 100 + w, c=c)
    raw = blob_oracle.raw_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], h, w, np.float64)
    _, dref = blob_oracle.composite(raw)
    want_d = np.moveaxis(dref, -1, 1)
    b = _blob(syn)
    feats = _cuda(syn["features"]).to(dtype)
    want_g = blob_oracle.splat_features_from_scores(want_d, _np(feats).astype(np.float64), None, channels_last=False)
    comp, grid = ops.render_fused(b["xs"], b["ys"], b["covs"], b["sizes"], feats, h, w)
    close_scaled(_np(comp), want_d, rel, f"composed {h}x{w}")
    close_scaled(_np(grid), want_g, rel, f"grid {h}x{w} C={c}")
    # same call into a buffer offset by one element: grid base not 8/16-byte aligned
    store = torch.zeros(n * c * h * w + 1, dtype=dtype, device=DEV)
    grid2 = store[1:].view(n, c, h, w)
    comp2 = torch.empty_like(comp)
    ops.render_fused_into(b["xs"], b["ys"], b["covs"], b["sizes"], feats, h, w, comp2, grid2)
    assert torch.equal(grid2, grid) and torch.equal(comp2, comp)
    assert float(store[0]) == 0.0


@pytest.mark.parametrize("n,k,levels,dtype,rel", [
    (3, 33, [(32, 32, 640), (16, 16, 1280), (8, 8, 1280)], torch.bfloat16, 1e-2),        # cfg3's lower levels
    (2, 33, [(64, 64, 320), (32, 32, 640), (16, 16, 1280), (8, 8, 1280)], torch.float32, 1e-5),
    (5, 17, [(12, 20, 320), (6, 10, 640), (3, 5, 320)], torch.float16, 2e-3),            # odd strides: generic path per unit
    (1, 65, [(16, 16, 640), (8, 8, 320)], torch.float32, 1e-5),
    (2, 33, [(16, 16, 96), (8, 8, 320)], torch.float32, 1e-5),                           # different channel tiles: level by level
    (2, 5, [(16, 16, 64), (8, 8, 3)], torch.float32, 1e-5)])                             # tiny K / C: FMA engine per level
def test_feature_splat_levels_vs_oracle(n, k, levels, dtype, rel):
    """blobsplat_feature_splat_levels: a pyramid of stage-3 problems (one launch when the levels share a tiling) against
    the float64 oracle per level, and bit-identical to the single-level entry point."""
    from blobctrl_b200 import ops
    g = torch.Generator().manual_seed(n * 131 + k)
    scs, fts = [], []
    for (h, w, c) in levels:
        sc = torch.rand(n, k, h, w, generator=g)
        scs.append((sc / sc.sum(1, keepdim=True)).to(DEV).to(dtype))
        fts.append(torch.randn(n, k, c, generator=g).to(DEV).to(dtype))
    outs = ops.feature_splat_levels(scs, fts)
    assert len(outs) == len(levels)
    fused = None
    if k >= 12 and all(c % 32 == 0 and c >= 64 for _, _, c in levels):
        fused = ops.feature_splat_levels(scs, fts, engine="tensor")      # the thread-staged single-launch path
    # 16-bit pyramids take the TMA engine under AUTO (splat_tma.cu): same products, another fp32 summation order, so the
    # engines may differ in the last 16-bit place — each is held to the oracle; float32 paths stay bit-identical
    exact = dtype == torch.float32
    for i, ((h, w, c), sc, ft, out) in enumerate(zip(levels, scs, fts, outs)):
        assert out.shape == (n, c, h, w) and out.dtype == dtype and out.is_contiguous()
        want = blob_oracle.splat_features_from_scores(_np(sc).astype(np.float64), _np(ft).astype(np.float64), None,
                                                      channels_last=False)
        close_scaled(_np(out), want, rel, f"level {h}x{w} C={c}")
        single = ops.feature_splat(sc, ft)
        close_scaled(_np(single), want, rel, f"single-level call, level {h}x{w} C={c}")
        if fused is not None:
            close_scaled(_np(fused[i]), want, rel, f"tensor engine, level {h}x{w} C={c}")
            assert torch.equal(fused[i], single), "single-launch pyramid (tensor engine) differs from the per-level launches"
        if exact:
            assert torch.equal(out, single), f"level {h}x{w}: differs from the single-level call"
    assert ops.feature_splat_levels([], []) == []


@pytest.mark.parametrize("n,m,h,w,c,dtype,rel", [
    (500, 12, 16, 16, 64, torch.bfloat16, 1e-2),     # 4 whole runs per CTA, 4 B buffers: staging warps + ring
    (700, 20, 16, 24, 96, torch.float16, 2e-3),      # 5 runs per CTA, 3 tiles each
    (460, 9, 16, 16, 64, torch.float32, 1e-5),       # float maps with a small B: the ring kernel in 3xTF32
    (300, 40, 20, 20, 320, torch.bfloat16, 1e-2),    # partial last tile (400 px), 320 channels
    (150, 12, 32, 32, 64, torch.bfloat16, 1e-2),     # 1-2 runs per CTA: single-buffer kernel for comparison
    (37, 127, 12, 12, 32, torch.float32, 1e-5),      # the blob limit of the fused kernel
    (40, 127, 16, 32, 96, torch.bfloat16, 1e-2),     # two-pixel kernel at the blob limit (Kp = 144, narrow drain sub-steps)
    (9, 33, 24, 40, 352, torch.bfloat16, 1e-2),      # two-pixel kernel: ragged second channel chunk, partial last tile
    (64, 32, 64, 64, 320, torch.float16, 2e-3),      # cfg3's level 64 in f16: equal tile ranges, partial runs
    (6, 16, 30, 31, 64, torch.bfloat16, 1e-2),       # odd width: stays on the one-pixel kernel
    (5, 20, 16, 16, 100, torch.float32, 1e-5),       # channel count off the 32-grid: ragged tile, scalar feature loads
    (3, 33, 32, 32, 1029, torch.bfloat16, 1e-2),     # 1029 channels: four chunks, the last one 69 wide
    (4, 12, 20, 24, 67, torch.float16, 2e-3),
    (5, 0, 9, 9, 32, torch.float32, 1e-5)])          # background only
def test_fused_render_schedules_and_staging_ring(n, m, h, w, c, dtype, rel):
    """blobsplat_render across its schedules (whole runs / equal ranges) and both staging schemes (compute warps with one
    B buffer / staging warps with a ring), every image against the float64 oracle."""
    from blobctrl_b200 import ops
    if m == 0:   # no blobs: the background takes every pixel (alpha 1, nothing in front)
        rng = np.random.default_rng(7)
        syn = {"xs": np.zeros((n, 0), np.float32), "ys": np.zeros((n, 0), np.float32), "sizes": np.zeros((n, 0), np.float32),
               "covs": np.zeros((n, 0, 2, 2), np.float32), "features": rng.standard_normal((n, 1, c)).astype(np.float32)}
        want_d = np.ones((n, 1, h, w))
    else:
        syn = blob_oracle.synthetic_blobs(n, m, seed=n + m, c=c)
        raw = blob_oracle.raw_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], h, w, np.float64)
        _, dref = blob_oracle.composite(raw)
        want_d = np.moveaxis(dref, -1, 1)
    b = _blob(syn)
    feats = _cuda(syn["features"]).to(dtype)
    want_g = blob_oracle.splat_features_from_scores(want_d, _np(feats).astype(np.float64), None, channels_last=False)
    comp, grid = ops.render_fused(b["xs"], b["ys"], b["covs"], b["sizes"], feats, h, w)
    close_scaled(_np(comp), want_d, rel, f"composed N={n} M={m}")
    if m == 0:
        return
    close_scaled(_np(grid), want_g, rel, f"grid N={n} M={m} C={c}")
    # the stand-alone stage 3 on the same weights takes the other A source through the same schedules
    g2 = ops.feature_splat(comp, feats, engine="tensor")
    want_g2 = blob_oracle.splat_features_from_scores(_np(comp).astype(np.float64), _np(feats).astype(np.float64), None,
                                                     channels_last=False)
    close_scaled(_np(g2), want_g2, rel, f"stage 3 from maps N={n} M={m} C={c}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_render_multiscale_one_call_equals_the_call_sequence(dtype):
    """blobsplat_render_multiscale == blobsplat_render + blobsplat_pyramid + blobsplat_feature_splat_levels, bit for bit,
    including a level that is wanted as maps only."""
    from blobctrl_b200 import ops
    n, m, s = 5, 20, 32
    syn = blob_oracle.synthetic_blobs(n, m, seed=77, c=1)
    b = _blob(syn)
    g = torch.Generator().manual_seed(3)
    feats = [torch.randn(n, m + 1, c, generator=g).to(DEV).to(dtype) if c else None for c in (64, 0, 160, 96)]
    comps, grids = ops.render_multiscale(b["xs"], b["ys"], b["covs"], b["sizes"], s, feats, dtype)
    d0, g0 = ops.render_fused(b["xs"], b["ys"], b["covs"], b["sizes"], feats[0], s, s, out_dtype=dtype)
    pyr = ops.halving_pyramid(d0, s >> 3)
    assert torch.equal(comps[0], d0) and torch.equal(grids[0], g0) and grids[1] is None
    lv = [l for l in range(1, 4) if feats[l] is not None]
    seq = ops.feature_splat_levels([pyr[s >> l] for l in lv], [feats[l] for l in lv])
    for l in range(1, 4):
        assert torch.equal(comps[l], pyr[s >> l])
    for l, want in zip(lv, seq):
        assert torch.equal(grids[l], want)


# ---------------------------------------------------------------------------------------------------------------------
# round 2: preview kernel, graph paths, error propagation, pipeline-method fixtures, N1 kernels, the cfg4 loop
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,rel", [(torch.float32, 1e-5), (torch.float64, 1e-9)])
def test_preview_one_launch_vs_oracle_and_the_three_launch_path(dtype, rel):
    """blobsplat_preview (stages 1+2 + colour splat in one launch) == utils.py:198-223 + :244-270 on the oracle, and the
    same image as the scores -> feature-splat path it replaces; multi-blob, per-image colours, H != W, gated blobs."""
    from blobctrl_b200 import ops
    U = _impl()
    for (n, m, h, w, seed) in [(1, 1, 64, 64, 1), (3, 7, 20, 36, 2), (2, 28, 17, 9, 3), (2, 150, 8, 8, 4)]:
        syn = blob_oracle.synthetic_blobs(n, m, seed=seed)
        b = _blob(syn, dtype)
        rng = np.random.default_rng(seed)
        colors = rng.random((n, m + 1, 3)) if m > 28 or seed == 2 else blob_oracle.BLOB_VIS_COLORS.astype(np.float64)
        raw = blob_oracle.raw_scores(_np(b["xs"]), _np(b["ys"]), _np(b["covs"]), syn["sizes"], h, w, np.float64)
        _, d = blob_oracle.composite(raw)                                  # [N,H,W,K]
        col = colors if colors.ndim == 3 else np.broadcast_to(colors[: m + 1], (n, m + 1, 3))
        cq = _np(_cuda(np.ascontiguousarray(col)).to(dtype)).astype(np.float64)
        want = np.einsum("nhwk,nkc->nchw", d, cq[:, : m + 1])
        img, comp = ops.render_preview(b["xs"], b["ys"], b["covs"], b["sizes"], _cuda(colors).to(dtype), h, w, want_composed=True)
        assert img.shape == (n, 3, h, w) and img.dtype == dtype
        close_scaled(_np(img), want, rel, f"preview {n}x{m}x{h}x{w}")
        close_scaled(_np(comp), np.moveaxis(d, -1, 1), rel, "preview composed")
        # the 8-bit picture (blobsplat_preview_u8): exactly the app's conversion of the float image (blobctrl_app.py:645-646),
        # permute and truncation done in the launch; within one count of the float64 oracle's picture
        u8 = ops.render_preview_u8(b["xs"], b["ys"], b["covs"], b["sizes"], _cuda(colors).to(dtype), h, w)
        assert u8.shape == (n, h, w, 3) and u8.dtype == torch.uint8
        app = (img.permute(0, 2, 3, 1).contiguous().cpu().numpy() * 255).astype(np.uint8)
        assert np.array_equal(u8.cpu().numpy(), app), f"u8 preview {n}x{m}x{h}x{w}"
        ref8 = (np.moveaxis(want, 1, -1) * 255).astype(np.uint8)
        assert np.abs(u8.cpu().numpy().astype(np.int16) - ref8.astype(np.int16)).max() <= 1
    # through the reference-signature API: the UI call (blobctrl_app.py:637-646) takes the one-launch path
    blob = blob_oracle.blob_from_ellipse(G.ellipses()[12]["ellipse"], 512, 512)
    bb = {k: _cuda(v).to(dtype) if k != "sizes" else _cuda(v) for k, v in blob.items()}
    got = U.get_blob_vis_img_from_blob_dict(bb, viz_size=(128, 128))
    d2, _ = ops.render_scores(bb["xs"], bb["ys"], bb["covs"], bb["sizes"], 128, 128)
    three = ops.feature_splat(d2, U.BLOB_VIS_COLORS[:2][None].to(DEV).to(dtype), engine="fma")
    assert got.dtype == dtype and (got - three).abs().max().item() <= (1e-6 if dtype == torch.float32 else 1e-14)
    # ... and as the picture itself, from device and from host-built dicts
    pic = U.get_blob_vis_u8_from_blob_dict(bb, viz_size=(128, 128))
    assert pic.shape == (128, 128, 3) and pic.dtype == np.uint8
    assert np.array_equal(pic, (got[0].permute(1, 2, 0).contiguous().cpu().numpy() * 255).astype(np.uint8))
    host = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in blob.items()}
    if dtype == torch.float64:
        assert np.array_equal(U.get_blob_vis_u8_from_blob_dict(host, viz_size=(128, 128)), pic)


def test_graphed_preview_picture_host_to_host():
    """The UI's preview step as one graph replay: host parameters in, host uint8 picture out (blobctrl_app.py:637-648);
    equal to the app's conversion of the eager float32 preview, call after call."""
    from blobctrl_b200.preview import preview_renderer
    U = _impl()
    r = preview_renderer((96, 96), DEV, picture=True)
    for i in (1, 12, 30, 5, 20):
        b = blob_oracle.blob_from_ellipse(G.ellipses()[i]["ellipse"], 512, 512)
        pic = r.render_picture(b["xs"], b["ys"], b["covs"]).copy()
        b32 = {k: _cuda(v).float() for k, v in b.items()}
        img = U.get_blob_vis_img_from_blob_dict(b32, viz_size=(96, 96))
        assert pic.shape == (96, 96, 3) and pic.dtype == np.uint8
        assert np.array_equal(pic, (img[0].permute(1, 2, 0).contiguous().cpu().numpy() * 255).astype(np.uint8)), i


def test_graphed_preview_back_to_back_without_sync():
    """consecutive GraphedBlobRenderer calls with no synchronisation between them must each render their own
    parameters (the pinned staging block is double-buffered behind an event)."""
    from blobctrl_b200.preview import preview_renderer
    U = _impl()
    r = preview_renderer((96, 96), DEV)
    idxs = [1, 12, 30, 5, 20, 33, 8]
    blobs = [blob_oracle.blob_from_ellipse(G.ellipses()[i]["ellipse"], 512, 512) for i in idxs]
    snaps = []
    for b in blobs:                                        # no sync between calls; snapshot in stream order
        _, img = r(b["xs"], b["ys"], b["covs"])
        snaps.append(img.clone())
    torch.cuda.synchronize()
    for b, img in zip(blobs, snaps):
        b32 = {k: _cuda(v).float() for k, v in b.items()}
        want = U.get_blob_vis_img_from_blob_dict(b32, viz_size=(96, 96))
        assert (img - want).abs().max().item() <= 2e-6


def test_cuda_graph_kwarg_replays_and_tracks_in_place_updates():
    """splat_features(..., cuda_graph=True): same values as the eager call; a second call with the same tensors is a replay
    that sees in-place parameter updates; a different shape captures its own graph."""
    from blobctrl_b200 import graphs
    U = _impl()
    graphs.clear()
    syn = blob_oracle.synthetic_blobs(1, 16, seed=3, c=320)
    b = _blob(syn); f = _cuda(syn["features"])
    eager = U.splat_features(**b, features=f, score_size=64, interp_size=64, ret_layout=False)
    g1 = U.splat_features(**b, features=f, score_size=64, interp_size=64, ret_layout=False, cuda_graph=True)
    assert torch.equal(g1["feature_grid"], eager["feature_grid"]) and torch.equal(g1["scores_pyramid"][64], eager["scores_pyramid"][64])
    syn2 = blob_oracle.synthetic_blobs(1, 16, seed=4, c=320)
    for k in ("xs", "ys", "covs", "sizes"):
        b[k].copy_(_cuda(syn2[k]))
    f.copy_(_cuda(syn2["features"]))
    g2 = U.splat_features(**b, features=f, score_size=64, interp_size=64, ret_layout=False, cuda_graph=True)
    eager2 = U.splat_features(**b, features=f, score_size=64, interp_size=64, ret_layout=False)
    assert g2["feature_grid"].data_ptr() == g1["feature_grid"].data_ptr()          # a replay into the static buffers
    assert torch.equal(g2["feature_grid"], eager2["feature_grid"])
    one = blob_oracle.blob_from_ellipse(G.ellipses()[3]["ellipse"], 512, 512)
    o32 = {k: _cuda(v).float() for k, v in one.items()}
    d = U.splat_features(**o32, score_size=(128, 128), return_d_score=True, cuda_graph=True)
    assert torch.equal(d, U.splat_features(**o32, score_size=(128, 128), return_d_score=True))
    graphs.clear()


def test_cuda_failure_propagates_unsupported_falls_back(monkeypatch):
    """a -3 status (CUDA failure) from the fused render must reach the caller; only -2 (outside the
    kernel's envelope, nothing launched) selects another engine."""
    from blobctrl_b200 import _capi, ops
    U = _impl()
    syn = blob_oracle.synthetic_blobs(2, 16, seed=5, c=64)
    b = _blob(syn); f = _cuda(syn["features"])
    want = U.splat_features(**b, features=f, score_size=16, interp_size=16, ret_layout=False, engine="fma")["feature_grid"]

    def boom(*a, **k):
        raise _capi.BlobSplatCudaError("blobsplat: CUDA failure (status -3): injected")
    monkeypatch.setattr(ops, "render_fused", boom)
    monkeypatch.setattr(ops, "render_multiscale", boom)
    with pytest.raises(_capi.BlobSplatCudaError):
        U.splat_features(**b, features=f, score_size=16, interp_size=16, ret_layout=False)
    with pytest.raises(_capi.BlobSplatCudaError):
        U.splat_features_multiscale(**b, score_size=16, level_features={16: f, 8: f})

    def unsupported(*a, **k):
        raise _capi.BlobSplatUnsupported("blobsplat: unsupported: injected")
    monkeypatch.setattr(ops, "render_fused", unsupported)
    monkeypatch.setattr(ops, "render_multiscale", unsupported)
    ms = U.splat_features_multiscale(**b, score_size=16, level_features={16: f, 8: f})
    assert (ms["feature_grids"][16] - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    got = U.splat_features(**b, features=f, score_size=16, interp_size=16, ret_layout=False)["feature_grid"]
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    # the real library reports -2 for a shape outside the tensor kernel (no launch), and the message says so
    with pytest.raises(_capi.BlobSplatUnsupported, match="unsupported"):
        ops.feature_splat(torch.rand(1, 130, 8, 8, device=DEV), torch.randn(1, 130, 64, device=DEV), engine="tensor")


def test_pipeline_method_fixtures_on_the_cuda_path():
    """a10 against fixtures recorded from the reference's own StableDiffusionBlobNetPipeline methods
    (tests/golden/make_golden_pipeline.py; pipeline_blobnet.py:706-739, :973-984), float16 and float32."""
    import json
    import os
    from blobctrl_b200.pipelines import (BlobConditioningMixin, BlobNetInputBuffers, construct_blobnet_input,
                                         prepare_blob_conditioning)
    z = np.load(os.path.join(G.GOLDEN, "pipeline.npz"))
    mix = BlobConditioningMixin()
    for c in json.load(open(os.path.join(G.GOLDEN, "pipeline_cases.json")))["cases"]:
        n = c["name"]
        if c["func"] == "splat_features_from_scores":
            sc, ft = _cuda(z[f"{n}/scores"]), _cuda(z[f"{n}/features"])
            got = mix.splat_features_from_scores(sc, ft, c["size"], channels_last=c["channels_last"])
            want = z[f"{n}/out"]
            assert got.dtype == sc.dtype and got.is_contiguous()
            rel = 2e-3 if sc.dtype == torch.float16 else 1e-5
            close_scaled(_np(got[:, ::c["c_stride"]]), want.astype(np.float64), rel, n)
            if sc.shape[1 if not c["channels_last"] else 3] == 1 and sc.dtype == torch.float16:
                # K = 1 in 16 bits: one exact product, one rounding -> bit-identical to the reference's einsum
                assert np.array_equal(_np(got[:, ::c["c_stride"]]), want.astype(np.float32)), n
        elif c["func"] == "construct_blobnet_input":
            t = {k: _cuda(z[f"{n}/{k}"]) for k in ("lat", "img", "scores", "feats", "fg", "bg")}
            assert torch.equal(mix.construct_blobnet_input(t["lat"], t["scores"], t["img"], t["feats"]), t["fg"])
            assert torch.equal(construct_blobnet_input(t["lat"], t["scores"], t["img"], background=True), t["bg"])
        else:                                              # the prologue :973-984 + one step's canvases
            gs = _cuda(z[f"{n}/gs_score"]); dino = _cuda(z[f"{n}/dino"])
            dt = dino.dtype
            cond = prepare_blob_conditioning(gs, dino, batch=c["batch"], dtype=dt, device=DEV)
            want_f = z[f"{n}/fg_gs_feats"]
            if dt == torch.float16:
                assert np.array_equal(_np(cond.fg_gs_feats), want_f.astype(np.float32)), n
            else:
                close_scaled(_np(cond.fg_gs_feats), want_f.astype(np.float64), 1e-5, n)
            lat, fg_lat, bg_lat = (_cuda(z[f"{n}/{k}"]) for k in ("lat", "fg_lat", "bg_lat"))
            x = construct_blobnet_input(lat, cond.fg_gs_scores, fg_lat, cond.fg_gs_feats)
            xb = construct_blobnet_input(lat, cond.bg_gs_scores, bg_lat, background=True)
            want_x, want_xb = _cuda(z[f"{n}/blobnet_model_input"]), _cuda(z[f"{n}/unet_bg_input"])
            assert torch.equal(xb, want_xb)
            bufs = BlobNetInputBuffers(c["batch"], lat.shape[2], lat.shape[3], dino.shape[-1], dt, DEV)
            bufs.fill_static(cond.fg_gs_scores, cond.bg_gs_scores, dino.repeat(c["batch"], 1, 1), fg_lat, bg_lat)
            px, pxb = bufs.update(lat)
            assert torch.equal(pxb, want_xb) and torch.equal(px, x)
            if dt == torch.float16:
                assert torch.equal(x, want_x), n
            else:
                close_scaled(_np(x), _np(want_x).astype(np.float64), 1e-5, n)


@pytest.mark.parametrize("dtype,rel", [(torch.float32, 1e-5), (torch.float16, 2e-3), (torch.bfloat16, 1e-2)])
def test_hoisted_conv_in_kernels_equal_the_full_convolution(dtype, rel):
    """N1 on the GPU: blobsplat_conv_in_weights + blobsplat_conv_in_hoisted == conv_in over the (4 + 1 + C)-plane canvas of
    construct_blobnet_input (models/blobnet.py:241-245, :840; pipeline_blobnet.py:724-739), K = 1 as the pipeline runs it
    and rank-K conditioning, against a float64 convolution of the reference canvas."""
    from blobctrl_b200.pipelines import HoistedConvIn, construct_blobnet_input
    g = torch.Generator().manual_seed(0)
    for (b2, h, w, c, o, k) in [(2, 12, 12, 20, 8, 1), (3, 16, 16, 1024, 320, 1), (2, 9, 20, 33, 40, 3), (1, 64, 64, 64, 64, 5)]:
        weight = (torch.randn(o, 4 + 1 + c, 3, 3, generator=g) * (9 * (5 + c)) ** -0.5).to(DEV).to(dtype)
        bias = torch.randn(o, generator=g).to(DEV).to(dtype)
        score = torch.rand(b2, 1, h, w, generator=g).to(DEV).to(dtype)
        sk = score if k == 1 else torch.rand(b2, k, h, w, generator=g).to(DEV).to(dtype)
        f = torch.randn(b2, k, c, generator=g).to(DEV).to(dtype)
        img_lat = torch.randn(b2, 4, h, w, generator=g).to(DEV).to(dtype)
        hoist = HoistedConvIn(weight, bias)
        hoist.prepare(score, sk, f)
        feats64 = torch.einsum("nkhw,nkc->nchw", sk.double(), f.double())
        for _ in range(2):
            lat = torch.randn(b2, 4, h, w, generator=g).to(DEV).to(dtype)
            canvas = construct_blobnet_input(lat.double(), score.double(), img_lat.double(), feats64)
            want = torch.nn.functional.conv2d(canvas, weight.double(), bias.double(), padding=1)
            got = hoist(torch.cat([img_lat, lat], dim=-1))
            assert got.shape == want.shape == (b2, o, h, 2 * w) and got.dtype == dtype
            close_scaled(_np(got), _np(want), rel, f"hoisted conv_in B={b2} C={c} O={o} K={k}")
    # the layer as the reference runs it (full canvas in, library convolution): same module, untouched semantics
    full = construct_blobnet_input(lat, score, img_lat, torch.einsum("nkhw,nkc->nchw", sk.float(), f.float()).to(dtype))
    assert hoist(full).shape == (b2, o, h, 2 * w)


def test_residual_injection_accepts_the_pipelines_cropped_views():
    """pipeline_blobnet.py:1085-1087 passes ``residual[..., -h:]`` (a strided view): consumed in place, same bits."""
    from blobctrl_b200.pipelines import inject_residual
    g = torch.Generator().manual_seed(12)
    for dtype in (torch.float16, torch.float32):
        hidden = torch.randn(3, 10, 8, 16, generator=g).to(DEV).to(dtype)
        res = torch.randn(3, 10, 8, 16, generator=g).to(DEV).to(dtype)
        crop = res[..., -8:]
        assert not crop.is_contiguous()
        want = hidden.clone(); want[..., -8:] = want[..., -8:] + crop
        assert torch.equal(inject_residual(hidden.clone(), crop, 1.0), want)
        odd = res[:, :, :, 3:11]                           # not the right-most columns: falls back to a contiguous copy
        want2 = hidden.clone(); want2[..., -8:] = want2[..., -8:] + odd
        assert torch.equal(inject_residual(hidden.clone(), odd, 1.0), want2)


@pytest.mark.skipif(not _have_reference(), reason="baseline/_ref not installed (scripts/install_reference.sh)")
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_cfg4_edit_loop_latent_parity_toy_width(dtype):
    """BASELINE configs[3] at toy width: the UNMODIFIED reference pipeline loop (pipeline_blobnet.py:1024-1102) vs the same
    object with the CUDA splat, persistent canvases (N2), hoisted conv_in (N1) and hooked residual injection (N4).
    N2 + N4 are bit-exact substitutions; N1 re-associates one convolution (<= 1e-2 of the latent scale in 16 bits)."""
    from baseline import cfg4_harness as H
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    r = H.compare_arms(device=DEV, dtype=dtype, batch=2, steps=6, small=True, warm_steps=1)
    assert np.isfinite(r["latent_absmax"]) and r["gs_score_max_abs_diff"] <= 1e-5
    a, b = r["ours_n2_n4"], r["ours_n1_n2_n4"]
    # toy UNet: conv_in + (1 pair + downsampler) + 1 resnet | mid | (2 resnets + upsampler) + 2 pairs = 10 residuals per step
    assert a["calls"]["injections"] == 6 * 10 and a["calls"]["splat_calls"] == 1 and a["calls"]["canvas_fills"] == 2
    assert a["calls"]["canvas_updates"] == 12
    assert a["latent_rel_to_absmax"] <= (1e-4 if dtype == torch.float32 else 1e-2), a
    assert b["latent_rel_to_absmax"] <= (1e-3 if dtype == torch.float32 else 1e-2), b


@pytest.mark.parametrize("dtype,rel", [(torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
@pytest.mark.parametrize("n,k,levels", [
    (2, 33, [(32, 32, 640), (16, 16, 1280), (8, 8, 1280)]),                    # cfg3's lower levels: one launch
    (3, 17, [(64, 64, 320)]),                                                  # 2.5 channel groups: clipped third group
    (1, 1, [(64, 64, 1024)]),                                                  # the pipeline's K = 1 splat: exact
    (2, 65, [(64, 64, 320), (32, 32, 640)]),                                   # Kp = 80: shallower rings
    (2, 5, [(24, 24, 72), (12, 12, 136), (4, 4, 8)]),                          # ragged pixel and channel tails, 16-pixel image
    (1, 128, [(48, 40, 200)]),                                                 # K = 128 (the thread-staged engine stops at 127 blobs + bg)
    (1, 256, [(16, 16, 64), (8, 8, 136)]),                                     # K = 256: the TMA box limit (16 MMA k-steps)
    (5, 33, [(8, 8, 64), (4, 2, 1280)]),                                       # 8-pixel image: one 16-column MMA
    (2, 40, [(20, 18, 328), (10, 12, 96), (64, 64, 320), (2, 4, 64)])])        # four levels of unrelated shapes
def test_tma_engine_vs_oracle(n, k, levels, dtype, rel):
    """splat_tma.cu (engine='tma'): operands by TMA, tcgen05 from shared memory, 128-byte line stores — every level against
    the float64 contraction of the same 16-bit inputs (utils.py:57-77), as one launch and level by level."""
    from blobctrl_b200 import ops
    g = torch.Generator().manual_seed(n * 17 + k)
    scs, fts = [], []
    for (h, w, c) in levels:
        sc = torch.rand(n, k, h, w, generator=g)
        scs.append((sc / sc.sum(1, keepdim=True)).to(DEV).to(dtype))
        fts.append(torch.randn(n, k, c, generator=g).to(DEV).to(dtype))
    outs = ops.feature_splat_levels(scs, fts, engine="tma")
    for (h, w, c), sc, ft, out in zip(levels, scs, fts, outs):
        assert out.shape == (n, c, h, w) and out.dtype == dtype and out.is_contiguous()
        want = torch.einsum("nkhw,nkc->nchw", sc.double(), ft.double()).cpu().numpy()
        close_scaled(_np(out), want, rel, f"level {h}x{w} C={c}")
        assert torch.equal(out, ops.feature_splat(sc, ft, engine="tma")), "one launch != level by level on the same engine"
        if k == 1:
            assert torch.equal(out, torch.einsum("nkhw,nkc->nchw", sc.float(), ft.float()).to(dtype)), "a single product is exact"


def test_tma_engine_views_and_envelope():
    """Strided score views are consumed in place; shapes outside the TMA envelope raise BlobSplatUnsupported on request and
    fall back to the other engines under AUTO; empty batches are no-ops."""
    from blobctrl_b200 import ops, _capi as C
    g = torch.Generator().manual_seed(5)
    empty = ops.feature_splat_levels([torch.empty(0, 17, 8, 8, device=DEV, dtype=torch.bfloat16)] * 2,
                                     [torch.empty(0, 17, 64, device=DEV, dtype=torch.bfloat16)] * 2, engine="tma")
    assert [tuple(t.shape) for t in empty] == [(0, 64, 8, 8)] * 2
    with pytest.raises(C.BlobSplatUnsupported):        # K = 257 > one TMA box
        ops.feature_splat(torch.rand(1, 257, 8, 8, device=DEV).bfloat16(), torch.randn(1, 257, 64, device=DEV).bfloat16(), engine="tma")
    big = torch.rand(2, 20, 40, 32, generator=g).to(DEV).to(torch.bfloat16)
    view = big[:, 2:19, :32, :]                          # plane stride 1280 > 1024 pixels, image stride 20 planes
    ft = torch.randn(2, 17, 256, generator=g).to(DEV).to(torch.bfloat16)
    out = ops.feature_splat(view, ft, engine="tma")
    want = torch.einsum("nkhw,nkc->nchw", view.double(), ft.double()).cpu().numpy()
    close_scaled(_np(out), want, 1e-2, "strided view")
    for sc, f in ((torch.rand(2, 17, 8, 8, device=DEV), torch.randn(2, 17, 64, device=DEV)),                       # float32
                  (torch.rand(2, 17, 5, 5, device=DEV).bfloat16(), torch.randn(2, 17, 64, device=DEV).bfloat16()),  # 25 pixels
                  (torch.rand(2, 17, 8, 8, device=DEV).bfloat16(), torch.randn(2, 17, 60, device=DEV).bfloat16())):  # C % 8 != 0
        with pytest.raises(C.BlobSplatUnsupported):
            ops.feature_splat(sc, f, engine="tma")
        ref = torch.einsum("nkhw,nkc->nchw", sc.double(), f.double()).cpu().numpy()
        close_scaled(_np(ops.feature_splat(sc, f)), ref, 1e-5 if sc.dtype == torch.float32 else 1e-2, "AUTO fallback")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,m,c,levels", [(5, 32, 320, 4), (3, 5, 64, 4), (2, 0, 32, 3), (70, 12, 96, 4), (4, 64, 352, 2),
                                          (3, 100, 640, 4)])
def test_fused_pyramid_equals_the_pyramid_kernel(n, m, c, levels, dtype):
    """64 x 64 multi-scale renders in 16 bits: the halving pyramid (utils.py:280-294) leaves the render launch itself
    (render_tc2.cuh kPyr: levels 1-2 per tile from staged half-sums, level 3 by the second tile of a pair).  Bit-identical to
    blobsplat_render + blobsplat_pyramid, on repeated calls (the arrival counters reset themselves) and under a CUDA graph."""
    from blobctrl_b200 import ops
    syn = blob_oracle.synthetic_blobs(n, m, seed=n + m, c=1)
    b = _blob(syn)
    g = torch.Generator().manual_seed(11)
    feats = [torch.randn(n, m + 1, c, generator=g).to(DEV).to(dtype)] + [None] * (levels - 1)
    d0, g0 = ops.render_fused(b["xs"], b["ys"], b["covs"], b["sizes"], feats[0], 64, 64, out_dtype=dtype)
    pyr = ops.halving_pyramid(d0, 64 >> (levels - 1))
    for rep in range(3):
        comps, grids = ops.render_multiscale(b["xs"], b["ys"], b["covs"], b["sizes"], 64, feats, dtype)
        assert torch.equal(comps[0], d0) and torch.equal(grids[0], g0)
        for l in range(1, levels):
            assert torch.equal(comps[l], pyr[64 >> l]), f"call {rep}: level {64 >> l} differs from the pyramid kernel"
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            comps, grids = ops.render_multiscale(b["xs"], b["ys"], b["covs"], b["sizes"], 64, feats, dtype)
        for _ in range(2):
            for t in comps:
                t.zero_()
            gr.replay()
    torch.cuda.synchronize()
    for l in range(1, levels):
        assert torch.equal(comps[l], pyr[64 >> l]), f"graph replay: level {64 >> l}"


@pytest.mark.parametrize("m,h,w,c", [(16, 64, 64, 320), (1, 64, 64, 1024), (32, 20, 36, 100), (8, 9, 7, 3), (0, 16, 16, 32),
                                     (5, 33, 1, 65), (24, 128, 128, 48)])
def test_one_image_latency_render_vs_oracle(m, h, w, c):
    """blobsplat_render_small — the CUDA-core kernel single small float32 renders take under AUTO (BASELINE config 2): composed
    maps within a last place of the stand-alone stages 1+2, grid within 1e-5 of the float64 oracle (full-fp32 products, so in fact
    ~1e-7), ragged pixel / channel tiles, gated blobs, K = 1; and the reference-signature call routes to it."""
    from blobctrl_b200 import ops
    U = _impl()
    if m == 0:                                                             # no blobs: the background owns every pixel
        z = torch.zeros(1, 0, device=DEV)
        feats = torch.randn(1, 1, c, generator=torch.Generator().manual_seed(5)).to(DEV)
        comp, grid = ops.render_small(z, z, torch.zeros(1, 0, 2, 2, device=DEV), z, feats, h, w)
        assert torch.equal(comp, torch.ones(1, 1, h, w, device=DEV))
        assert torch.equal(grid, feats[0, 0].view(1, c, 1, 1).expand(1, c, h, w))
        return
    syn = blob_oracle.synthetic_blobs(1, m, seed=31 + m, c=c)
    if m >= 5:
        syn["sizes"][0, 2] = 0.0                                           # a gated blob (utils.py:165-172)
    b = _blob(syn)
    feats = torch.from_numpy(syn["features"]).to(DEV)
    comp, grid = ops.render_small(b["xs"], b["ys"], b["covs"], b["sizes"], feats, h, w)
    sc, _ = ops.render_scores(b["xs"], b["ys"], b["covs"], b["sizes"], h, w)
    assert (comp - sc).abs().max().item() <= 5e-7          # same coefficients; the 4-pixel kernel steps u, v along x (last-place differences)
    raw = blob_oracle.raw_scores(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], h, w, np.float64)
    _, d = blob_oracle.composite(raw)
    want = np.einsum("nhwk,nkc->nchw", d, syn["features"].astype(np.float64))
    close_scaled(_np(comp), np.moveaxis(d, -1, 1), 1e-5, "latency render: composed")
    close_scaled(_np(grid), want, 1e-5, "latency render: grid")
    assert np.abs(_np(grid) - want).max() <= 2e-6 * max(np.abs(want).max(), 1e-30)      # full fp32: far inside the bar
    if h == w and ops.small_render_applies(1, m, h, w, c):
        out = U.splat_features(**b, features=feats, score_size=h, interp_size=h, ret_layout=False)
        assert torch.equal(out["feature_grid"], grid) and torch.equal(out["scores_pyramid"][h], comp)
        forced = U.splat_features(**b, features=feats, score_size=h, interp_size=h, ret_layout=False, engine="tensor") if c >= 8 and m >= 1 else None
        if forced is not None:
            close_scaled(_np(forced["feature_grid"]), want, 1e-5, "tensor engine on the same image")
    assert not ops.small_render_applies(2, m, h, w, c) and not ops.small_render_applies(1, 17, h, w, c)
    with pytest.raises(Exception):
        big = blob_oracle.synthetic_blobs(1, 40, seed=1, c=8)
        ops.render_small(*[_cuda(big[k]) for k in ("xs", "ys", "covs", "sizes")], _cuda(big["features"]), 8, 8)


def test_f32_exact_switch_keeps_float32_on_the_fma_engine(monkeypatch):
    """BLOBSPLAT_F32_EXACT=1: float32 stage 3 under AUTO runs on the FMA engine — bit-identical to
    engine='fma' — and splat_features gives the scores + FMA-splat result instead of the split-precision fused render."""
    from blobctrl_b200 import ops
    syn = blob_oracle.synthetic_blobs(3, 20, seed=9, c=96)
    b = _blob(syn)
    feats = torch.from_numpy(syn["features"]).to(DEV)
    sc, _ = ops.render_scores(b["xs"], b["ys"], b["covs"], b["sizes"], 32, 32)
    fma = ops.feature_splat(sc, feats, engine="fma")
    auto_default = ops.feature_splat(sc, feats)
    monkeypatch.setenv("BLOBSPLAT_F32_EXACT", "1")
    auto_exact = ops.feature_splat(sc, feats)
    assert torch.equal(auto_exact, fma)
    got = _impl().splat_features(**b, features=feats, score_size=32, interp_size=32, ret_layout=False)
    assert torch.equal(got["feature_grid"], fma)
    monkeypatch.delenv("BLOBSPLAT_F32_EXACT")
    want = blob_oracle.splat_features(**syn, score_size=32, interp_size=32, dtype=np.float64, ret_layout=False)
    close_scaled(_np(auto_default), want["feature_grid"], 1e-5, "default AUTO (tensor engine)")
    close_scaled(_np(fma), want["feature_grid"], 1e-5, "FMA engine")


@pytest.mark.parametrize("scale", [1.0, 3e-5, 0.3, 700.0, 5000.0, 2e7])
def test_fused_float32_render_across_feature_magnitudes(scale):
    """2xFP16 split of float32 maps: features of a unit are pre-scaled by a power of two only when max|f| lies outside
    [0.5, 4096) (render_tc.cuh, tc_stage_b) — both branches, and a mix of magnitudes across images, at the 1e-5 bar."""
    from blobctrl_b200 import ops
    n, m, c = 6, 40, 128
    syn = blob_oracle.synthetic_blobs(n, m, seed=21, c=c)
    per_image = np.array([scale, scale * 3.0, 1.0, scale, 1e-3, scale * 0.1], dtype=np.float32).reshape(n, 1, 1)
    syn["features"] = (syn["features"] * per_image).astype(np.float32)
    b = _blob(syn)
    feats = torch.from_numpy(syn["features"]).to(DEV)
    comp, grid = ops.render_fused(b["xs"], b["ys"], b["covs"], b["sizes"], feats, 32, 32)
    want = blob_oracle.splat_features(**syn, score_size=32, interp_size=32, dtype=np.float64, ret_layout=False)
    got, ref = _np(grid), want["feature_grid"]
    for i in range(n):                                  # per image: each has its own magnitude
        close_scaled(got[i], ref[i], 1e-5, f"image {i} (feature scale {float(per_image[i, 0, 0]):g})")


def test_write_combined_pinned_inputs_round_trip():
    """blobctrl_b200.hostmem: cudaHostAllocWriteCombined buffers are seen as pinned by torch, feed HostRenderer like
    ordinary pinned tensors (same maps, bit for bit) and are released explicitly."""
    from blobctrl_b200 import hostmem
    from blobctrl_b200.streaming import HostRenderer
    n, m, s, c = 8, 5, 32, 64
    syn = blob_oracle.synthetic_blobs(n, m, seed=77, c=c)
    host = {k: torch.from_numpy(v) for k, v in syn.items()}
    wc = hostmem.pinned_like(host, write_combined=True)
    pl = hostmem.pinned_like(host, write_combined=False)
    assert all(t.is_pinned() for t in wc.values()) and all(t.is_pinned() for t in pl.values())
    r = HostRenderer(n, m, s, c, torch.float32, DEV, chunks=2)
    a = r(wc["xs"], wc["ys"], wc["covs"], wc["sizes"], wc["features"])
    a = (a["scores_pyramid"][s].clone(), a["feature_grid"].clone())
    b = r(pl["xs"], pl["ys"], pl["covs"], pl["sizes"], pl["features"])
    torch.cuda.synchronize()
    assert torch.equal(a[0], b["scores_pyramid"][s]) and torch.equal(a[1], b["feature_grid"])
    for t in wc.values():
        hostmem.free_pinned(t)
