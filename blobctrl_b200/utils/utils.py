"""Drop-in for the renderer half of the reference's ``blobctrl/utils/utils.py``.

Same names, keyword arguments, defaults, return shapes, dict keys and error behaviour as the
reference functions cited below; the arithmetic runs in hand-written sm_100a CUDA kernels behind
the C ABI of ``include/blobsplat.h`` (see ``blobctrl_b200/ops.py``).  There is no CPU path and no PyTorch
fallback: the maps are always rendered by the CUDA kernels and always live on the GPU.  ``splat_features`` takes the
blob dict as the reference's scripts build it — host tensors (scripts/blobctrl_inference.py:101-109) — by uploading the
few bytes of parameters to the current CUDA device first; without a CUDA device it raises.

Behavioural notes (each mirrors a line of the reference):
  * pixel centres sit at integer coordinates, x = p % W, y = p // W, no half-pixel offset (utils.py:139-142);
  * ``sizes`` is an existence gate: < 0.5 -> score 1e-6 (utils.py:165-172);
  * channel 0 is the background (alpha 1), blob m is channel m+1, the highest index is front-most
    (utils.py:175-181);
  * a tuple ``score_size`` / ``viz_size`` only works for one image with one blob (utils.py:132-134,
    157-159) and raises RuntimeError otherwise, like the reference's ``.view(1, 1, H, W)``;
  * the maps keep the dtype of ``covs`` (float64 in -> float64 out).  Extension: ``out_dtype=`` asks for
    bfloat16/float16 maps from float32 parameters (the reference cannot run those dtypes at all).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from .. import _capi as C
from .. import ops

def _palette() -> Tensor:
    """utils.py:22-53 — the 29 preview colours (row 0 = background).  The reference stores them as
    4-decimal floats, kept exactly here as three channel tuples."""
    r = (0.9804, 1.0, 0.961, 0.8980, 0.3647, 0.3216, 0.6000, 0.1843, 0.6471, 0.8549, 0.4627, 0.8000, 0.9294,
         0.1412, 0.4000, 0.9647, 0.9725, 0.8627, 0.5294, 0.6196, 0.9961, 0.7882, 0.5451, 0.7059, 0.7020, 0.5216,
         0.8510, 0.6863, 0.4510)
    g = (0.9451, 0.494, 0.882, 0.5255, 0.4118, 0.7373, 0.7882, 0.5412, 0.6667, 0.6471, 0.3059, 0.3804, 0.3922,
         0.4745, 0.7725, 0.8118, 0.6118, 0.6902, 0.7725, 0.7255, 0.5333, 0.8588, 0.8784, 0.5922, 0.7020, 0.3608,
         0.6863, 0.3922, 0.4353)
    b = (0.9176, 0.357, 0.827, 0.0235, 0.6941, 0.6392, 0.2706, 0.7686, 0.6000, 0.1059, 0.6235, 0.6902, 0.3529,
         0.4235, 0.8000, 0.4431, 0.4549, 0.9490, 0.3725, 0.9529, 0.6941, 0.4549, 0.6431, 0.9059, 0.7020, 0.4588,
         0.4196, 0.3451, 0.298)
    return torch.tensor([r, g, b], dtype=torch.float32).t().contiguous()


BLOB_VIS_COLORS = _palette()


def viz_score_fn(score):
    """utils.py:387-390 — identity."""
    return score


_IDENTITY_VIZ_SCORE_FN = viz_score_fn


# --------------------------------------------------------------------------------------------------
# stage 3
# --------------------------------------------------------------------------------------------------
def splat_features_from_scores(scores: Tensor, features: Tensor, size: Optional[int],
                               channels_last: bool = True, engine: str = "auto") -> Tensor:
    """utils.py:57-77.

    Args:
        scores: [N, H, W, M] (or [N, M, H, W] if not channels_last)
        features: [N, M, C]; cast to the dtype/device of ``scores`` (utils.py:69)
        size: map size to return; when it differs from ``scores.shape[2]`` the scores are bilinearly
            resized first (align_corners=False)
    Returns: [N, C, H, W], contiguous
    """
    C.require_cuda(scores, "scores")
    if size and not (scores.shape[2] == size):
        nkhw = scores.permute(0, 3, 1, 2) if channels_last else scores
        oh, ow = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
        if (nkhw.shape[2], nkhw.shape[3]) != (oh, ow):
            nkhw = ops.resize_bilinear(nkhw, oh, ow)
        return ops.feature_splat(nkhw, features, channels_last=False, engine=engine)
    return ops.feature_splat(scores, features, channels_last=channels_last, engine=engine)


def pyramid_resize(img: Tensor, cutoff: int) -> Dict[int, Tensor]:
    """utils.py:280-294 — halve (bilinear) while the last dim exceeds ``cutoff``; dict keyed by last dim."""
    return ops.halving_pyramid(img, cutoff)


@torch.no_grad()
def visualize_features(viz_size=64, n_gaussians=None, scores=None, viz_colors=None) -> Dict[str, Tensor]:
    """utils.py:244-270 — stage 3 with colours as features.  ``scores`` is [N, H, W, K]."""
    k = n_gaussians + 1
    rand_colors = viz_colors is None
    viz_colors = viz_colors.to(scores.device) if not rand_colors else torch.rand((k, 3)).to(scores.device)
    if viz_colors.ndim == 2:
        viz_colors = viz_colors[:k][None].repeat_interleave(len(scores), 0)
    elif viz_colors.ndim == 3:
        viz_colors = viz_colors[:, :k]
    else:
        viz_colors = torch.rand((k, 3), device=scores.device)
    img = splat_features_from_scores(scores, viz_colors, viz_size)
    if rand_colors:
        imax = img.amax((2, 3))[:, :, None, None]
        imin = img.amin((2, 3))[:, :, None, None]
        img = img.sub(imin).div((imax - imin).clamp(min=1e-5)).mul(2).sub(1)
    return {"feature_img": img}


# --------------------------------------------------------------------------------------------------
# the renderer
# --------------------------------------------------------------------------------------------------
def _render_hw(covs: Tensor, score_size, viz_size) -> Tuple[int, int]:
    n_blobs = covs.shape[0] * covs.shape[1]
    if not isinstance(viz_size, int) and viz_size is not None:          # utils.py:120
        h, w = viz_size
        tuple_path = True
    elif isinstance(score_size, int):                                   # utils.py:137
        return score_size, score_size
    else:                                                               # utils.py:145
        h, w = score_size
        tuple_path = True
    if tuple_path and n_blobs != 1:
        raise RuntimeError(f"shape '[1, 1, {h}, {w}]' is invalid for input of size {n_blobs * h * w}: a tuple "
                           f"score_size/viz_size renders one image with one blob (reference utils.py:132-134,157-159)")
    return int(h), int(w)


def _viz_same_size(viz_size, h: int, w: int) -> bool:
    """visualize_features resizes the score maps to ``viz_size`` (utils.py:70-73); the one-launch preview applies when
    that resize is the identity, i.e. the maps were rendered at viz_size."""
    if isinstance(viz_size, int):
        return viz_size == h == w
    return viz_size is not None and (int(viz_size[0]), int(viz_size[1])) == (h, w)


def _upload_host_blobs(xs, ys, covs, sizes, features, kwargs):
    if not torch.cuda.is_available():
        C.require_cuda(covs, "covs")
    dev = torch.device("cuda", torch.cuda.current_device())
    up = lambda t: t.to(dev, non_blocking=True) if torch.is_tensor(t) and not t.is_cuda else t
    if torch.is_tensor(kwargs.get("viz_colors")):
        kwargs = dict(kwargs, viz_colors=up(kwargs["viz_colors"]))
    return up(xs), up(ys), up(covs), up(sizes), up(features), kwargs


def splat_features(
        xs: Tensor,
        ys: Tensor,
        covs: Tensor,
        sizes: Tensor,
        score_size: Optional[Union[int, Tuple[int, int]]] = None,
        interp_size: int = None,
        features: Tensor = None,
        viz_size: Optional[Union[int, Tuple[int, int]]] = None,
        is_viz: bool = False,
        ret_layout: bool = True,
        viz_score_fn=None,
        return_d_score=False,
        only_vis: bool = False,
        only_splatting_fg: bool = False,
        only_splatting_bg: bool = False,
        **kwargs) -> Union[Dict, Tensor]:
    """utils.py:80-241 — same arguments and returns.

    Args:
        xs, ys: [N, M] blob centres in [0, 1]         covs: [N, M, 2, 2]        sizes: [N, M] (or [N, M, 1])
        features: [N, M+1, C] (row 0 = background)    score_size: render size   interp_size: feature-grid size
    Returns:
        ``return_d_score``: composed maps [N, K | M | 1, H, W];
        ``only_vis``: {'feature_img': [N, 3, H, W]};
        else dict with scores_pyramid {size: [N,K,S,S]}, feature_grid [N,C,S,S], feature_img, entropy_img
        (+ xs, ys, covs, raw_scores [N,H,W,K], sizes, composed_scores [N,H,W,K], features when ret_layout).
    Extra keyword-only extensions (absent from the reference): ``out_dtype``, ``composite_mode``
    ('auto' | 'lane_pixel' | 'warp_scan'), ``engine`` ('auto' | 'fma' | 'tensor'), ``cuda_graph`` (bool, default False:
    replay a captured CUDA graph of this exact call — the latency path for N = 1; the returned tensors are then the
    graph's static buffers, see blobctrl_b200/graphs.py), ``check_singular`` (bool, default False: raise like the reference's
    ``torch.linalg.solve`` (utils.py:143) when a covariance has a zero or non-finite determinant — one device-to-host
    synchronisation per call, which is why it is opt-in; without it such a blob yields non-finite maps).
    """
    if kwargs.get("cuda_graph"):
        # opt-in: replay a captured graph of this exact call (blobctrl_b200/graphs.py); outputs are static buffers
        from .. import graphs
        kw = {k: v for k, v in kwargs.items() if k != "cuda_graph"}
        tensors = (xs, ys, covs, sizes, features, kw.get("viz_colors"))
        if graphs.capturable(*tensors) and (viz_score_fn is None or viz_score_fn is _IDENTITY_VIZ_SCORE_FN):
            key = ("splat_features", graphs.tensor_key(tensors), score_size, viz_size, interp_size, is_viz,
                   ret_layout, viz_score_fn is None, return_d_score, only_vis, only_splatting_fg, only_splatting_bg,
                   tuple(sorted((k, v) for k, v in kw.items() if not torch.is_tensor(v))) if len(kw) > 1 else
                   tuple((k, v) for k, v in kw.items() if not torch.is_tensor(v)))
            return graphs.call(key, lambda: splat_features(xs, ys, covs, sizes, score_size, interp_size, features, viz_size,
                                                           is_viz, ret_layout, viz_score_fn, return_d_score, only_vis,
                                                           only_splatting_fg, only_splatting_bg, **kw), keepalive=tensors)
    out_dtype = kwargs.get("out_dtype")
    composite_mode = kwargs.get("composite_mode", "auto")
    engine = kwargs.get("engine", "auto")
    if not covs.is_cuda:
        # the scripts' blob dict (blobctrl_inference.py:101-109, blobctrl_app.py:520-528) is built from numpy on the host and
        # the result moved with .to(device) / .cpu() afterwards (:174, blobctrl_app.py:645): upload the parameters, render on
        # the GPU, return GPU maps.  No device -> require_cuda raises: nothing is ever computed on the host.
        xs, ys, covs, sizes, features, kwargs = _upload_host_blobs(xs, ys, covs, sizes, features, kwargs)
    if kwargs.get("check_singular"):
        c64 = covs.double()
        det = c64[..., 0, 0] * c64[..., 1, 1] - c64[..., 0, 1] * c64[..., 1, 0]
        if not bool(((det != 0) & torch.isfinite(det)).all()):          # the reference's error text (torch.linalg.solve)
            raise torch.linalg.LinAlgError("torch.linalg.solve: The solver failed because the input matrix is singular.")
    h, w = _render_hw(covs, score_size, viz_size)
    m = covs.shape[1]
    select = "bg" if only_splatting_bg else ("fg" if only_splatting_fg else "all")

    if return_d_score:                                                  # utils.py:193-194
        d, _ = ops.render_scores(xs, ys, covs, sizes, h, w, select=select, out_dtype=out_dtype,
                                 composite_mode=composite_mode)
        return d

    wants_grid = not only_vis
    viz_colors = kwargs.get("viz_colors", None)
    if (only_vis and is_viz and select == "all" and torch.is_tensor(viz_colors) and viz_colors.ndim in (2, 3)
            and (viz_score_fn is None or viz_score_fn is _IDENTITY_VIZ_SCORE_FN) and out_dtype is None
            and covs.dtype in (torch.float32, torch.float64) and _viz_same_size(viz_size, h, w)
            and viz_colors.shape[-2] >= m + 1 and (viz_colors.ndim == 2 or viz_colors.shape[0] == covs.shape[0])):
        # the UI preview (blobctrl_app.py:637-650): stages 1+2 and the colour splat in ONE launch, no score maps in HBM
        img, _ = ops.render_preview(xs, ys, covs, sizes, viz_colors, h, w)
        return {"feature_img": img}
    if wants_grid and interp_size is None:                              # utils.py:291 — `W > None`
        raise TypeError("'>' not supported between instances of 'int' and 'NoneType' (interp_size is required "
                        "unless return_d_score or only_vis)")
    custom_fn = is_viz and viz_score_fn is not None and viz_score_fn is not _IDENTITY_VIZ_SCORE_FN
    need_raw = (wants_grid and ret_layout) or custom_fn

    # one fused launch when the grid is wanted at render resolution and nothing needs the raw scores
    d = raw = grid = None
    if (wants_grid and not need_raw and not is_viz and select == "all" and features is not None
            and isinstance(score_size, int) and interp_size == score_size and engine != "fma"
            and covs.dtype != torch.float64):
        try:
            if (engine == "auto" and covs.dtype == torch.float32 and features.dtype == torch.float32
                    and out_dtype in (None, torch.float32) and not C.f32_exact()
                    and ops.small_render_applies(covs.shape[0], m, h, w, features.shape[-1])):
                # ONE small image (the scripts' and the UI's case, BASELINE config 2): the CUDA-core latency kernel
                d, grid = ops.render_small(xs, ys, covs, sizes, features, h, w)
            else:
                d, grid = ops.render_fused(xs, ys, covs, sizes, features, h, w, out_dtype=out_dtype or covs.dtype)
        except C.BlobSplatUnsupported:
            if engine == "tensor":
                raise
            d = grid = None                                             # outside the tensor kernel's envelope
    if d is None:
        d, raw = ops.render_scores(xs, ys, covs, sizes, h, w, select=select, want_raw=need_raw,
                                   out_dtype=out_dtype, composite_mode=composite_mode)

    ret = {}
    if is_viz:                                                          # utils.py:198-214
        if viz_score_fn is None:
            scores_viz = d.permute(0, 2, 3, 1)
        elif not custom_fn:                                             # identity: composite(raw) == d_all
            d_all = d if select == "all" else ops.render_scores(xs, ys, covs, sizes, h, w, out_dtype=out_dtype)[0]
            scores_viz = d_all.permute(0, 2, 3, 1)
        else:
            posterior = viz_score_fn(raw.permute(0, 2, 3, 1))
            scores_viz = ops.composite(posterior.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        ret.update(visualize_features(viz_size, m, scores_viz, kwargs.get("viz_colors", None)))  # utils.py:213-214
    if only_vis:                                                        # utils.py:222-223
        return ret

    ret["scores_pyramid"] = pyramid_resize(d, cutoff=interp_size)       # utils.py:226-227
    if grid is None:
        grid = splat_features_from_scores(ret["scores_pyramid"][interp_size], features, interp_size,
                                          channels_last=False, engine=engine)
    ret.update({"feature_grid": grid, "feature_img": None, "entropy_img": None})
    if ret_layout:                                                      # utils.py:236-239
        sz = sizes.squeeze(-1) if torch.is_tensor(sizes) and sizes.ndim == 3 else sizes
        ret.update({"xs": xs, "ys": ys, "covs": covs, "raw_scores": raw.permute(0, 2, 3, 1), "sizes": sz,
                    "composed_scores": d.permute(0, 2, 3, 1), "features": features})
    return ret


def splat_features_multiscale(xs: Tensor, ys: Tensor, covs: Tensor, sizes: Tensor, score_size: int,
                              level_features: Dict[int, Tensor], out_dtype: Optional[torch.dtype] = None,
                              engine: str = "auto") -> Dict[str, Dict[int, Tensor]]:
    """BlobNet multi-scale conditioning (BASELINE config 3; SURVEY.md §0.1 — a synthetic extension built
    from the reference's own pieces): render at ``score_size``, ``pyramid_resize`` down to the smallest
    requested level and ``splat_features_from_scores`` the per-level features [N, M+1, C_S] at every
    level S (C_S = BlobNet block_out_channels, models/blobnet.py:172).

    Returns {'scores_pyramid': {S: [N,K,S,S]}, 'feature_grids': {S: [N,C_S,S,S]}}.
    """
    if not level_features:
        raise ValueError("level_features is empty")
    # the common shape — consecutive halvings from score_size down, features at the top level, 16/32-bit maps — is one
    # C call (blobsplat_render_multiscale): at small batches the per-call host overhead bounds the multi-call path
    top_f = level_features.get(score_size)
    n_lv = 1
    while (score_size >> n_lv) >= min(level_features) and (score_size % (1 << n_lv)) == 0:
        n_lv += 1
    dt = out_dtype or covs.dtype
    if (engine == "auto" and top_f is not None and n_lv <= 4 and covs.dtype != torch.float64 and dt != torch.float64
            and all(s in [score_size >> l for l in range(n_lv)] for s in level_features)):
        try:
            comps, gl = ops.render_multiscale(xs, ys, covs, sizes, score_size,
                                              [level_features.get(score_size >> l) for l in range(n_lv)], dt)
            return {"scores_pyramid": {score_size >> l: comps[l] for l in range(n_lv)},
                    "feature_grids": {score_size >> l: gl[l] for l in range(n_lv) if gl[l] is not None}}
        except C.BlobSplatUnsupported:
            pass                                     # outside the fused render's envelope: the general path below
    grids = {}
    d = None
    top = level_features.get(score_size)
    if top is not None and engine != "fma" and covs.dtype != torch.float64:
        # the full-resolution level comes out of the fused render (stages 1+2+3 in one launch)
        try:
            d, grids[score_size] = ops.render_fused(xs, ys, covs, sizes, top, score_size, score_size,
                                                    out_dtype=out_dtype or covs.dtype)
        except C.BlobSplatUnsupported:
            if engine == "tensor":
                raise
            d = None
    if d is None:
        d, _ = ops.render_scores(xs, ys, covs, sizes, score_size, score_size, out_dtype=out_dtype)
    pyr = pyramid_resize(d, cutoff=min(level_features))
    rest = [s for s in level_features if s not in grids]
    # the remaining levels: one launch for the whole pyramid where the levels allow it (blobsplat_feature_splat_levels)
    for s, g in zip(rest, ops.feature_splat_levels([pyr[s] for s in rest], [level_features[s].to(d.dtype) for s in rest],
                                                   engine=engine)):
        grids[s] = g
    return {"scores_pyramid": pyr, "feature_grids": grids}


def splat_ellipses(ellipses, sizes=None, image_size=(512, 512), score_size=64, only_splatting_fg=False,
                   only_splatting_bg=False, out_dtype: torch.dtype = torch.float32) -> Tensor:
    """Device-side ellipse front end: the whole host recipe of the reference's callers —
    get_gs_from_ellipse -> normalize_gs -> get_blob_dict_from_norm_gs -> splat_features(return_d_score=True)
    (scripts/blobctrl_inference.py:71-117) — for a batch, in one launch.

    ellipses: [N, M, 5] tensor (or nested list) of (xc, yc, d1, d2, angle_deg), cv2.fitEllipse convention, in pixels
    of an ``image_size`` = (height, width) image; blob index = depth order (highest in front).
    score_size: int S or (H, W).  Returns composed maps [N, M+1 | M | 1, H, W] (float32 / bfloat16 / float16).
    """
    e = torch.as_tensor(ellipses, dtype=torch.float32)
    if e.ndim == 1:
        e = e.view(1, 1, 5)
    elif e.ndim == 2:
        e = e.unsqueeze(0)
    if not e.is_cuda:
        e = e.cuda()                       # a few floats per blob: the host->device hop of the parameters
    h, w = (score_size, score_size) if isinstance(score_size, int) else score_size
    select = "bg" if only_splatting_bg else ("fg" if only_splatting_fg else "all")
    return ops.render_scores_from_ellipses(e, sizes, image_size, int(h), int(w), select=select, out_dtype=out_dtype)[0]


def flatten_cv_ellipse(ellipse):
    """((xc, yc), (d1, d2), angle) -> [xc, yc, d1, d2, angle] (the layout splat_ellipses takes)."""
    (xc, yc), (d1, d2), ang = ellipse
    return [float(xc), float(yc), float(d1), float(d2), float(ang)]


# --------------------------------------------------------------------------------------------------
# geometry (host side, float64) — the input contract of the renderer
# --------------------------------------------------------------------------------------------------
def rotation_matrix(theta: Tensor) -> Tensor:
    """utils.py:273-276."""
    cos, sin = torch.cos(theta), torch.sin(theta)
    return torch.stack([cos, sin, -sin, cos], dim=-1).view(*theta.shape, 2, 2)


def ellipse_to_gaussian(x, y, a, b, theta):
    """utils.py:297-341: mean (x, y); Sigma = R(theta) diag(b^2, a^2) R(theta)^T with negated off-diagonals
    (a = minor semi-axis, b = major semi-axis, theta = CCW angle of the major axis, radians)."""
    c, s = np.cos(theta), np.sin(theta)
    rot = np.array([[c, -s], [s, c]])
    cov = rot @ np.array([[b ** 2, 0], [0, a ** 2]]) @ rot.T
    cov[0, 1] *= -1
    cov[1, 0] *= -1
    return np.array([x, y]), cov


def gaussian_to_ellipse(mean, cov_matrix):
    """utils.py:344-384: inverse of ellipse_to_gaussian up to the reference's angle convention."""
    x, y = mean
    evals, evecs = np.linalg.eig(cov_matrix)
    b, a = np.sqrt(max(evals)), np.sqrt(min(evals))
    v = evecs[:, np.argmin(evals)]
    ang = np.degrees(np.arctan2(v[1], v[0]))
    if ang < 0:
        ang += 180
    return x, y, a, b, ang


def get_theta_anti_clockwise_long_axis(angle_clockwise_short_axis):
    """scripts/blobctrl_inference.py:71-75."""
    return np.radians((((180 - angle_clockwise_short_axis) % 180) + 90) % 180)


def get_gs_from_ellipse(ellipse):
    """scripts/blobctrl_inference.py:78-85: OpenCV ((xc,yc),(d1,d2),angle) -> (mean_px, cov_px)."""
    (xc, yc), (d1, d2), ang = ellipse
    return ellipse_to_gaussian(xc, yc, d1 / 2, d2 / 2, get_theta_anti_clockwise_long_axis(ang))


def normalize_gs(mean, cov_matrix_rotated, width, height):
    """scripts/blobctrl_inference.py:88-98."""
    max_length = np.sqrt(width ** 2 + height ** 2)
    return mean / np.array([width, height]), cov_matrix_rotated / (max_length ** 2)


def get_blob_dict_from_norm_gs(normalized_mean, normalized_cov_matrix, device="cuda"):
    """scripts/blobctrl_inference.py:101-109, with the tensors created on ``device``."""
    xs, ys = normalized_mean
    return {"xs": torch.tensor(xs, device=device).unsqueeze(0), "ys": torch.tensor(ys, device=device).unsqueeze(0),
            "covs": torch.tensor(normalized_cov_matrix, device=device).unsqueeze(0).unsqueeze(0),
            "sizes": torch.tensor([1.0], device=device).unsqueeze(0)}


def get_blob_score_from_blob_dict(blob, score_size=64):
    """scripts/blobctrl_inference.py:112-117."""
    return splat_features(**blob, score_size=score_size, return_d_score=True)[0]


def get_blob_vis_img_from_blob_dict(blob, viz_size=64, score_size=64):
    """scripts/blobctrl_app.py:637-646 up to the tensor (the caller converts to PIL)."""
    return splat_features(**blob, interp_size=64, viz_size=viz_size, is_viz=True, ret_layout=True,
                          score_size=score_size, viz_score_fn=viz_score_fn, viz_colors=BLOB_VIS_COLORS,
                          only_vis=True)["feature_img"]


def get_blob_vis_u8_from_blob_dict(blob, viz_size=64):
    """scripts/blobctrl_app.py:637-648 up to ``Image.fromarray``: the preview of image 0 as a host uint8 array [H, W, 3]
    (``Image.fromarray(get_blob_vis_u8_from_blob_dict(blob, viz_size))`` is the app's ``blob_vis_img``).  One launch
    renders, permutes and converts; the device-to-host copy moves 3 bytes per pixel."""
    covs = blob["covs"]
    if not covs.is_cuda:
        xs, ys, covs, sizes, _, _ = _upload_host_blobs(blob["xs"], blob["ys"], covs, blob["sizes"], None, {})
    else:
        xs, ys, sizes = blob["xs"], blob["ys"], blob["sizes"]
    h, w = (viz_size, viz_size) if isinstance(viz_size, int) else (int(viz_size[0]), int(viz_size[1]))
    if isinstance(viz_size, tuple) and (covs.shape[0] != 1 or covs.shape[1] != 1):
        raise RuntimeError(f"shape '[1, 1, {h}, {w}]' is invalid for input of size {covs.shape[0] * covs.shape[1] * h * w}")
    return ops.render_preview_u8(xs, ys, covs, sizes, BLOB_VIS_COLORS, h, w)[0].cpu().numpy()


# --------------------------------------------------------------------------------------------------
# overlay helpers (host side, OpenCV) — same names and behaviour as utils.py:393-456; imported by the scripts
# (scripts/blobctrl_app.py:26)
# --------------------------------------------------------------------------------------------------
def vis_scores(blob_d_score, viz_size):
    """utils.py:393-402.  Like the reference it loads the colour table from ``BED_CONF_COLORS.pt`` in the working
    directory (the reference ships no such file, so its own call raises FileNotFoundError too)."""
    n_gaussians = blob_d_score.shape[1] - 1
    scores_viz = blob_d_score.permute(0, 2, 3, 1)
    viz_colors = torch.load("BED_CONF_COLORS.pt")
    return visualize_features(viz_size=viz_size, n_gaussians=n_gaussians, scores=scores_viz, viz_colors=viz_colors)["feature_img"]


def vis_gt_ellipse_from_norm_gs(img, gt_mus, gt_covs, color=None):
    """utils.py:405-430 — img: [h, w, c] uint8; gt_mus: [n, 2], gt_covs: [n, 2, 2] normalised gaussians.  Returns a copy."""
    import cv2
    result = img.copy()
    height, width, _ = img.shape
    max_length = np.sqrt(width ** 2 + height ** 2)
    for mu, cov in zip(gt_mus, gt_covs):
        mean = tuple(v.numpy() for v in mu.cpu().unbind(-1))
        xc, yc, a, b, angle = gaussian_to_ellipse(mean, cov.cpu().numpy())
        ellipse = (xc * width, yc * height), (a * max_length * 2, b * max_length * 2), angle
        if color is None:
            color = [255, 0, 0]
        cv2.ellipse(result, ellipse, color, 3)
    return result


def vis_gt_ellipse_from_norm_ellipse(img, norm_ellipse, color=None):
    """utils.py:433-444 — draws in place on ``img`` and returns it."""
    import cv2
    height, width, _ = img.shape
    max_length = np.sqrt(width ** 2 + height ** 2)
    (xc, yc), (d1, d2), theta = norm_ellipse
    ellipse = (xc * width, yc * height), (d1 * max_length, d2 * max_length), theta
    cv2.ellipse(img, ellipse, [255, 0, 0] if color is None else color, 3)
    return img


def vis_gt_ellipse_from_ellipse(img, ellipse, color=None):
    """utils.py:447-456 — draws the OpenCV ellipse ((xc, yc), (d1, d2), angle) in place on ``img`` and returns it."""
    import cv2
    (xc, yc), (d1, d2), theta = ellipse
    cv2.ellipse(img, ((xc, yc), (d1, d2), theta), [255, 0, 0] if color is None else color, 3)
    return img
