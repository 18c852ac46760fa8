"""CPU: host-side logic — sharding (incl. a world_size-2 gloo run), conditioning layout, input
canonicalisation errors, geometry recipe, and the rule that the product never touches oracle/."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import golden_util as G
from oracle import blob_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    from blobctrl_b200.sharding import shard_bounds
    for n in (0, 1, 7, 8, 1024, 1027):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_shard_blobs_slices_per_image_tensors():
    from blobctrl_b200.sharding import shard_blobs
    syn = {k: torch.from_numpy(v) for k, v in blob_oracle.synthetic_blobs(5, 3, seed=1, c=4).items()}
    parts = [shard_blobs(syn, r, 2) for r in range(2)]
    assert parts[0]["covs"].shape[0] == 3 and parts[1]["covs"].shape[0] == 2
    for k in syn:
        assert torch.equal(torch.cat([p[k] for p in parts]), syn[k])


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from blobctrl_b200.sharding import shard_bounds, gather_maps
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
r = dist.get_rank()
N = 5
full = torch.arange(N * 2 * 3, dtype=torch.float32).reshape(N, 2, 3)
lo, hi = shard_bounds(N, r, 2)
out = gather_maps(full[lo:hi].clone(), N)
assert torch.equal(out, full), (r, out)
even = gather_maps(full[r * 2:(r + 1) * 2].clone(), 4)
assert torch.equal(even, full[:4])
dist.barrier(); dist.destroy_process_group()
print("ok", r)
"""


def test_gather_maps_world_size_2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_construct_blobnet_input_layout_cpu():
    """pipeline_blobnet.py:724-739 — pure layout; runs on CPU tensors."""
    from blobctrl_b200.pipelines import BlobConditioningMixin, construct_blobnet_input
    lat, img = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    sc, ft = torch.rand(2, 1, 8, 8), torch.randn(2, 6, 8, 8)
    x = construct_blobnet_input(lat, sc, img, ft)
    assert x.shape == (2, 11, 8, 16)
    assert torch.equal(x[..., :8], torch.cat([img, sc, ft], 1)) and torch.equal(x[..., 8:], torch.cat([lat, sc, ft], 1))
    xb = BlobConditioningMixin().construct_blobnet_input(lat, sc, img, background=True)
    assert xb.shape == (2, 5, 8, 16) and torch.equal(xb[:, :4, :, 8:], lat)


def test_cpu_tensors_fail_loudly():
    import blobctrl_b200 as B
    syn = {k: torch.from_numpy(v) for k, v in blob_oracle.synthetic_blobs(1, 2, seed=1, c=4).items()}
    with pytest.raises(RuntimeError, match="no CPU"):
        B.splat_features(syn["xs"], syn["ys"], syn["covs"], syn["sizes"], score_size=8, return_d_score=True)
    with pytest.raises(RuntimeError, match="no CPU"):
        B.splat_features_from_scores(torch.rand(1, 3, 4, 4), torch.rand(1, 3, 2), 4, channels_last=False)
    with pytest.raises(RuntimeError, match="no CPU"):
        B.pyramid_resize(torch.rand(1, 3, 8, 8), 4)


def test_geometry_recipe_matches_oracle_and_reference_fixture():
    """ellipse -> gaussian -> normalised blob dict (blobctrl_inference.py:71-109)."""
    import blobctrl_b200.utils.utils as U
    for e in G.ellipses():
        mean, cov = U.get_gs_from_ellipse(e["ellipse"])
        nm, nc = U.normalize_gs(mean, cov, 512, 512)
        ob = blob_oracle.blob_from_ellipse(e["ellipse"], 512, 512)
        assert np.array_equal(nc, ob["covs"][0, 0]) and nm[0] == ob["xs"][0] and nm[1] == ob["ys"][0]
        assert abs(nc[0, 1] - nc[1, 0]) <= 1e-15 * abs(nc).max()
    # round trip through gaussian_to_ellipse on a well-conditioned ellipse
    mean, cov = U.ellipse_to_gaussian(10.0, 20.0, 3.0, 7.0, 0.3)
    x, y, a, b, _ = U.gaussian_to_ellipse(mean, cov)
    assert (x, y) == (10.0, 20.0) and abs(a - 3.0) < 1e-9 and abs(b - 7.0) < 1e-9
    r = U.rotation_matrix(torch.tensor([0.3]))
    assert torch.allclose(r[0] @ r[0].T, torch.eye(2), atol=1e-6)
    assert torch.equal(U.BLOB_VIS_COLORS, torch.from_numpy(blob_oracle.BLOB_VIS_COLORS))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under blobctrl_b200/ may import, call or read it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[./]", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "blobctrl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} references oracle/"
    assert "blobctrl_b200" in sys.modules or True
    import blobctrl_b200  # noqa: F401
    assert not any(m == "oracle" or m.startswith("oracle.") for m in sys.modules
                   if "blobctrl_b200" in (getattr(sys.modules[m], "__file__", "") or ""))


def test_hoisted_conv_in_has_no_cpu_path():
    """N1 runs on the CUDA kernels only (tests/test_gpu_parity.py checks it against the full convolution)."""
    from blobctrl_b200.pipelines import HoistedConvIn
    hoist = HoistedConvIn(torch.randn(8, 4 + 1 + 6, 3, 3), torch.randn(8))
    with pytest.raises(RuntimeError, match="no CPU"):
        hoist.prepare(torch.rand(2, 1, 4, 4), torch.rand(2, 1, 4, 4), torch.randn(2, 1, 6))
    full = hoist(torch.randn(2, 11, 4, 8))                 # a full canvas still goes through the original layer
    assert full.shape == (2, 8, 4, 8)


def test_import_surface_matches_the_reference_package():
    """blobctrl/utils/__init__.py:1-2 exports splat_features, viz_score_fn, BLOB_VIS_COLORS, vis_gt_ellipse_from_ellipse;
    scripts/blobctrl_app.py:26 imports exactly those; the other helpers of utils.py:387-456 exist under the same names."""
    import blobctrl_b200.utils as pkg
    from blobctrl_b200.utils import BLOB_VIS_COLORS, splat_features, vis_gt_ellipse_from_ellipse, viz_score_fn  # noqa: F401
    for name in ("splat_features", "viz_score_fn", "BLOB_VIS_COLORS", "vis_gt_ellipse_from_ellipse"):
        assert name in pkg.__all__
    import blobctrl_b200.utils.utils as U
    for name in ("vis_scores", "vis_gt_ellipse_from_norm_gs", "vis_gt_ellipse_from_norm_ellipse", "vis_gt_ellipse_from_ellipse",
                 "ellipse_to_gaussian", "gaussian_to_ellipse", "rotation_matrix", "pyramid_resize", "visualize_features",
                 "splat_features_from_scores"):
        assert callable(getattr(U, name)), name
    # the overlay helpers draw with OpenCV exactly like utils.py:433-456: in place, 3 px, red by default
    img = np.zeros((64, 64, 3), np.uint8)
    out = U.vis_gt_ellipse_from_ellipse(img, ((32.0, 32.0), (20.0, 30.0), 15.0))
    assert out is img and img[..., 0].max() == 255 and img[..., 1:].max() == 0
    img2 = np.zeros((64, 64, 3), np.uint8)
    U.vis_gt_ellipse_from_norm_ellipse(img2, ((0.5, 0.5), (20 / np.sqrt(2 * 64 ** 2), 30 / np.sqrt(2 * 64 ** 2)), 15.0), color=[0, 255, 0])
    assert img2[..., 1].max() == 255 and np.array_equal(img2[..., 1] > 0, img[..., 0] > 0)
    mean, cov = U.ellipse_to_gaussian(0.5, 0.5, 0.1, 0.2, 0.3)
    img3 = np.zeros((64, 64, 3), np.uint8)
    res = U.vis_gt_ellipse_from_norm_gs(img3, torch.tensor([mean]), torch.tensor(np.array([cov])))
    assert res is not img3 and res[..., 0].max() == 255 and img3.max() == 0
    with pytest.raises(FileNotFoundError):              # utils.py:398 loads a file the reference does not ship
        U.vis_scores(torch.rand(1, 2, 4, 4), 4)


def test_status_codes_map_to_distinct_exceptions(monkeypatch):
    """-2 (unsupported: nothing ran) and -3 (CUDA failure: context may be poisoned) must not be confusable: only the
    former may be answered with another engine."""
    from blobctrl_b200 import _capi as C
    monkeypatch.setattr(C, "last_error", lambda: "boom")
    C.check(0)
    with pytest.raises(ValueError):
        C.check(-1)
    with pytest.raises(C.BlobSplatUnsupported):
        C.check(-2)
    with pytest.raises(C.BlobSplatCudaError):
        C.check(-3)
    assert issubclass(C.BlobSplatUnsupported, C.BlobSplatError) and issubclass(C.BlobSplatCudaError, C.BlobSplatError)
    assert not issubclass(C.BlobSplatCudaError, C.BlobSplatUnsupported)
    # no handler in the product catches the base class or a bare Exception around a kernel call
    pat = re.compile(r"except\s+(Exception|BaseException|C\.BlobSplatError|_capi\.BlobSplatError)\b|except\s*:")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "blobctrl_b200")):
        for f in files:
            if f.endswith(".py") and f != "_capi.py":
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f"{f} swallows CUDA failures"


def test_pipeline_method_fixtures_vs_oracle():
    """tests/golden/pipeline.npz (recorded from the reference's own pipeline methods, pipeline_blobnet.py:706-739) against
    the CPU oracle: pins the oracle for a10 before the GPU tests use it."""
    import json
    from oracle import aten_port
    z = np.load(os.path.join(ROOT, "tests", "golden", "pipeline.npz"))
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "pipeline_cases.json")))["cases"]
    assert len(cases) >= 10
    for c in cases:
        n = c["name"]
        if c["func"] == "splat_features_from_scores":
            sc, ft = torch.from_numpy(z[f"{n}/scores"]), torch.from_numpy(z[f"{n}/features"])
            got = aten_port.feature_splat(sc, ft, c["size"], c["channels_last"])[:, ::c["c_stride"]]
            assert torch.equal(got, torch.from_numpy(z[f"{n}/out"])), n
        elif c["func"] == "construct_blobnet_input":
            from blobctrl_b200.pipelines import construct_blobnet_input
            t = {k: torch.from_numpy(z[f"{n}/{k}"]) for k in ("lat", "img", "scores", "feats", "fg", "bg")}
            assert torch.equal(construct_blobnet_input(t["lat"], t["scores"], t["img"], t["feats"]), t["fg"])
            assert torch.equal(construct_blobnet_input(t["lat"], t["scores"], t["img"], background=True), t["bg"])


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "blobctrl")),
                    reason="baseline/_ref not installed (scripts/install_reference.sh)")
def test_cfg4_harness_stock_arm_runs_on_cpu():
    """The cfg4 harness drives the UNMODIFIED reference pipeline class: toy-width models, 2 steps, CPU, deterministic."""
    from baseline import cfg4_harness as H
    pipe = H.build_pipeline("cpu", torch.float32, small=True, feat_channels=64)
    assert type(pipe).__mro__[1].__module__ == "blobctrl.pipelines.pipeline_blobnet"
    assert pipe.blobnet.conv_in.in_channels == 4 + 1 + 64 and pipe.unet.conv_in.in_channels == 5
    inp = H.make_inputs(pipe, 1, "cpu", torch.float32, small=True)
    gs = H.reference_gs_score(pipe.unet.config.sample_size)
    assert gs.shape == (1, 2, 16, 16) and gs.dtype == torch.float64
    a, _ = H.run_edit(pipe, inp, gs, 2, "cpu", None)
    b, _ = H.run_edit(pipe, inp, gs, 2, "cpu", None)
    assert a.shape == (1, 4, 16, 16) and torch.isfinite(a).all() and torch.equal(a, b)
