"""TEST INFRASTRUCTURE ONLY — numpy restatement of BlobCtrl's blob renderer.

This is the CPU oracle the CUDA path is checked against.  It is a restatement of the
*algorithm* in ``/root/reference/blobctrl/utils/utils.py`` (cited per function as
``utils.py:<lines>``), written independently in numpy: closed-form 2x2 inverse instead of
``torch.linalg.solve``, an explicit back-to-front loop instead of flip/cumprod/roll.

Pinned by ``tests/test_oracle_golden.py`` against fixtures produced by running the real
reference in the build container (``tests/golden/make_golden.py``): agreement is ~1e-13
in float64.  The product package never imports this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

GATE_SCORE = np.float32(1e-6)  # utils.py:172 — a float32 scalar, promoted to the score dtype

# utils.py:22-53 (29 RGB rows).  Stored transposed/flattened so this table is data, not a copy
# of the reference's source layout.
_VIS_R = (0.9804, 1.0, 0.961, 0.8980, 0.3647, 0.3216, 0.6000, 0.1843, 0.6471, 0.8549, 0.4627, 0.8000,
          0.9294, 0.1412, 0.4000, 0.9647, 0.9725, 0.8627, 0.5294, 0.6196, 0.9961, 0.7882, 0.5451, 0.7059,
          0.7020, 0.5216, 0.8510, 0.6863, 0.4510)
_VIS_G = (0.9451, 0.494, 0.882, 0.5255, 0.4118, 0.7373, 0.7882, 0.5412, 0.6667, 0.6471, 0.3059, 0.3804,
          0.3922, 0.4745, 0.7725, 0.8118, 0.6118, 0.6902, 0.7725, 0.7255, 0.5333, 0.8588, 0.8784, 0.5922,
          0.7020, 0.3608, 0.6863, 0.3922, 0.4353)
_VIS_B = (0.9176, 0.357, 0.827, 0.0235, 0.6941, 0.6392, 0.2706, 0.7686, 0.6000, 0.1059, 0.6235, 0.6902,
          0.3529, 0.4235, 0.8000, 0.4431, 0.4549, 0.9490, 0.3725, 0.9529, 0.6941, 0.4549, 0.6431, 0.9059,
          0.7020, 0.4588, 0.4196, 0.3451, 0.298)
BLOB_VIS_COLORS = np.stack([_VIS_R, _VIS_G, _VIS_B], axis=1).astype(np.float32)  # [29, 3]


# --------------------------------------------------------------------------------------
# geometry (host side) — utils.py:297-341, scripts/blobctrl_inference.py:71-109
# --------------------------------------------------------------------------------------
def theta_from_cv_angle(angle_clockwise_short_axis: float) -> float:
    """blobctrl_inference.py:71-75: OpenCV fitEllipse angle -> CCW long-axis angle (rad)."""
    a = (180.0 - angle_clockwise_short_axis) % 180.0
    return float(np.radians((a + 90.0) % 180.0))


def ellipse_to_gaussian(x, y, a, b, theta) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:297-341: Sigma = R diag(b^2, a^2) R^T with the off-diagonals negated."""
    c, s = np.cos(theta), np.sin(theta)
    rot = np.array([[c, -s], [s, c]], dtype=np.float64)
    cov = rot @ np.array([[b ** 2, 0], [0, a ** 2]]) @ rot.T
    cov[0, 1] = -cov[0, 1]
    cov[1, 0] = -cov[1, 0]
    return np.array([x, y], dtype=np.float64), cov


def gs_from_ellipse(ellipse) -> Tuple[np.ndarray, np.ndarray]:
    """blobctrl_inference.py:78-85: ((xc,yc),(d1,d2),angle_deg) -> (mean_px, cov_px)."""
    (xc, yc), (d1, d2), ang = ellipse
    return ellipse_to_gaussian(xc, yc, d1 / 2.0, d2 / 2.0, theta_from_cv_angle(ang))


def normalize_gs(mean, cov, width, height) -> Tuple[np.ndarray, np.ndarray]:
    """blobctrl_inference.py:88-98: mean/(W,H); cov/(W^2+H^2)."""
    diag = np.sqrt(width ** 2 + height ** 2)
    return mean / np.array([width, height]), cov / (diag ** 2)


def blob_from_ellipse(ellipse, width=512, height=512) -> Dict[str, np.ndarray]:
    """blobctrl_inference.py:101-109: the blob dict every reference caller builds
    (xs, ys: shape [1]; covs [1,1,2,2] float64; sizes [[1.0]] float32)."""
    mean, cov = gs_from_ellipse(ellipse)
    nm, nc = normalize_gs(mean, cov, width, height)
    return {
        "xs": np.array([nm[0]], dtype=np.float64),
        "ys": np.array([nm[1]], dtype=np.float64),
        "covs": nc[None, None].astype(np.float64),
        "sizes": np.array([[1.0]], dtype=np.float32),
    }


# --------------------------------------------------------------------------------------
# stage 1 — utils.py:120-172
# --------------------------------------------------------------------------------------
def _canon(xs, ys, covs, sizes):
    covs = np.asarray(covs)
    n, m = covs.shape[0], covs.shape[1]
    xs = np.broadcast_to(np.asarray(xs).reshape(-1, m) if np.asarray(xs).size == n * m
                         else np.asarray(xs), (n, m))
    ys = np.broadcast_to(np.asarray(ys).reshape(-1, m) if np.asarray(ys).size == n * m
                         else np.asarray(ys), (n, m))
    sizes = np.asarray(sizes)
    if sizes.ndim == 3:  # utils.py:165-166
        sizes = sizes[..., 0]
    return xs, ys, covs, sizes, n, m


def raw_scores(xs, ys, covs, sizes, height: int, width: int, dtype=np.float64) -> np.ndarray:
    """Stages 1 + gate, returns [N, H, W, M] in ``dtype``.

    utils.py:138-143 (square) / :147-156 (tuple): delta = (pixel - centre*size)/size with
    pixel = (p % W, p // W); q = delta^T Sigma^-1 delta.  utils.py:162-163: s = min(1, 2*sigmoid(-q)).
    utils.py:165-172: sizes < 0.5 -> s := 1e-6.
    """
    xs, ys, covs, sizes, n, m = _canon(xs, ys, covs, sizes)
    dt = np.dtype(dtype)
    xs = xs.astype(dt); ys = ys.astype(dt); covs = covs.astype(dt)
    gx = np.arange(width, dtype=dt)[None, None, None, :]   # x = p % W
    gy = np.arange(height, dtype=dt)[None, None, :, None]  # y = p // W
    dx = (gx - (xs * dt.type(width))[:, :, None, None]) / dt.type(width)
    dy = (gy - (ys * dt.type(height))[:, :, None, None]) / dt.type(height)
    a = covs[:, :, 0, 0][:, :, None, None]
    b = covs[:, :, 0, 1][:, :, None, None]
    c = covs[:, :, 1, 0][:, :, None, None]
    d = covs[:, :, 1, 1][:, :, None, None]
    det = a * d - b * c
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        # solve(Sigma, delta) in closed form; q = delta . (Sigma^-1 delta)
        sx = (d * dx - b * dy) / det
        sy = (a * dy - c * dx) / det
        q = dx * sx + dy * sy
        s = dt.type(2) / (dt.type(1) + np.exp(q))
    s = np.where(np.isnan(s), dt.type(0), s)   # 0*inf at a degenerate blob's far pixels -> reference gives 0
    s = np.minimum(s, dt.type(1))
    gate = (sizes < 0.5)[:, :, None, None]
    s = np.where(gate, GATE_SCORE.astype(dt), s)
    return np.ascontiguousarray(np.moveaxis(s, 1, -1))  # [N,H,W,M]


# --------------------------------------------------------------------------------------
# stage 2 — utils.py:175-181
# --------------------------------------------------------------------------------------
def composite_with_bg(s: np.ndarray) -> np.ndarray:
    """s [N,H,W,K] (channel 0 = background alpha) -> composed d [N,H,W,K].

    utils.py:179-181: d_k = s_k * prod_{j>k}(1 - s_j), d_{K-1} = s_{K-1}.  Walk k = K-1..0 carrying
    the transmittance; emit before updating.  The carry is kept in float64 because ATen's CPU
    cumprod accumulates float32 inputs in double and rounds each output (SURVEY.md probe B10).
    """
    k = s.shape[-1]
    one = s.dtype.type(1)
    d = np.empty_like(s)
    trans = np.ones(s.shape[:-1], dtype=np.float64)
    for j in range(k - 1, -1, -1):
        d[..., j] = s[..., j] if j == k - 1 else trans.astype(s.dtype) * s[..., j]
        trans = trans * (one - s[..., j]).astype(np.float64)
    return d


def composite(raw: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """raw [N,H,W,M] -> (scores_with_bg [N,H,W,K], composed [N,H,W,K]), K = M+1.
    utils.py:175-176 prepends background alpha 1."""
    s = np.concatenate([np.ones_like(raw[..., :1]), raw], axis=-1)
    return s, composite_with_bg(s)


def render_scores(xs, ys, covs, sizes, height: int, width: int, dtype=np.float64,
                  select: str = "all") -> np.ndarray:
    """Composed maps [N, K|M|1, H, W] — utils.py:183-194 selection + 'n h w m -> n m h w'."""
    _, d = composite(raw_scores(xs, ys, covs, sizes, height, width, dtype))
    if select == "bg":
        d = d[..., :1]
    elif select == "fg":
        d = d[..., 1:]
    return np.ascontiguousarray(np.moveaxis(d, -1, 1))


# --------------------------------------------------------------------------------------
# bilinear resize / pyramid — utils.py:280-294, :70-73 (ATen upsample_bilinear2d, align_corners=False)
# --------------------------------------------------------------------------------------
def _src_index(out_size: int, in_size: int, dtype):
    scale = dtype(in_size) / dtype(out_size)
    src = (np.arange(out_size, dtype=dtype) + dtype(0.5)) * scale - dtype(0.5)
    src = np.maximum(src, dtype(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    lam1 = (src - i0.astype(dtype)).astype(dtype)
    return i0, i1, dtype(1) - lam1, lam1


def bilinear_resize(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """[N,C,H,W] -> [N,C,out_h,out_w]; matches F.interpolate(mode='bilinear', align_corners=False)."""
    dt = img.dtype.type
    h0, h1, hl0, hl1 = _src_index(out_h, img.shape[2], dt)
    w0, w1, wl0, wl1 = _src_index(out_w, img.shape[3], dt)
    top = img[:, :, h0][:, :, :, w0] * wl0 + img[:, :, h0][:, :, :, w1] * wl1
    bot = img[:, :, h1][:, :, :, w0] * wl0 + img[:, :, h1][:, :, :, w1] * wl1
    return (hl0[:, None] * top + hl1[:, None] * bot).astype(img.dtype)


def pyramid_resize(img: np.ndarray, cutoff: int) -> Dict[int, np.ndarray]:
    """utils.py:280-294: halve (bilinear) while last dim > cutoff; dict keyed by last dim."""
    out = [img]
    while img.shape[-1] > cutoff:
        s = img.shape[-1] // 2
        img = bilinear_resize(img, s, s)
        out.append(img)
    return {int(i.shape[-1]): i for i in out}


# --------------------------------------------------------------------------------------
# stage 3 — utils.py:57-77 (and pipeline_blobnet.py:706-721)
# --------------------------------------------------------------------------------------
def splat_features_from_scores(scores: np.ndarray, features: np.ndarray, size, channels_last=True):
    """out[n,c,h,w] = sum_m scores[n,m,h,w] * features[n,m,c]; resize first when size != H."""
    features = features.astype(scores.dtype)
    if size and not (scores.shape[2] == size):
        if channels_last:
            scores = np.moveaxis(scores, -1, 1)
        oh, ow = (size, size) if isinstance(size, int) else size
        scores = bilinear_resize(scores, oh, ow)
    elif channels_last:
        scores = np.moveaxis(scores, -1, 1)
    return np.ascontiguousarray(np.einsum("nmhw,nmc->nchw", scores, features))


def visualize_features(viz_size, n_gaussians, scores, viz_colors):
    """utils.py:244-270 with explicit colours (the random-colour branch is not reproducible)."""
    k = n_gaussians + 1
    vc = np.asarray(viz_colors)
    vc = np.repeat(vc[:k][None], len(scores), 0) if vc.ndim == 2 else vc[:, :k]
    return {"feature_img": splat_features_from_scores(scores, vc, viz_size)}


# --------------------------------------------------------------------------------------
# the whole renderer — utils.py:80-241
# --------------------------------------------------------------------------------------
def splat_features(xs, ys, covs, sizes, score_size=None, interp_size=None, features=None,
                   viz_size=None, is_viz=False, ret_layout=True, viz_score_fn=None,
                   return_d_score=False, only_vis=False, only_splatting_fg=False,
                   only_splatting_bg=False, dtype=None, **kwargs):
    covs = np.asarray(covs)
    dt = np.dtype(dtype) if dtype is not None else covs.dtype
    n, m = covs.shape[:2]
    if viz_size is not None and not isinstance(viz_size, int):        # utils.py:120 (N=M=1 only)
        h, w = viz_size
        if n * m != 1:
            raise RuntimeError("tuple viz_size path supports one image / one blob (utils.py:132-134)")
    elif isinstance(score_size, int):                                 # utils.py:137
        h = w = score_size
    else:                                                             # utils.py:145 (N=M=1 only)
        h, w = score_size
        if n * m != 1:
            raise RuntimeError("tuple score_size path supports one image / one blob (utils.py:157-159)")
    raw = raw_scores(xs, ys, covs, sizes, h, w, dt)
    scores, d = composite(raw)
    if only_splatting_bg:
        d = d[..., :1]
    elif only_splatting_fg:
        d = d[..., 1:]
    if return_d_score:
        return np.moveaxis(d, -1, 1)
    ret = {}
    if is_viz:
        sv = d
        if viz_score_fn is not None:                                  # utils.py:199-209
            sv = composite_with_bg(viz_score_fn(scores))
        ret.update(visualize_features(viz_size, m, sv, kwargs.get("viz_colors")))
    if only_vis:
        return ret
    score_img = np.moveaxis(d, -1, 1)
    ret["scores_pyramid"] = pyramid_resize(score_img, cutoff=interp_size)
    ret["feature_grid"] = splat_features_from_scores(ret["scores_pyramid"][interp_size], np.asarray(features),
                                                     interp_size, channels_last=False)
    ret.update({"feature_img": None, "entropy_img": None})
    if ret_layout:
        ret.update({"xs": xs, "ys": ys, "covs": covs, "raw_scores": scores, "sizes": sizes,
                    "composed_scores": d, "features": features})
    return ret


# --------------------------------------------------------------------------------------
# seeded synthetic blob sets — SURVEY.md §8(d)
# --------------------------------------------------------------------------------------
def synthetic_blobs(n: int, m: int, seed: int = 0, thin: bool = False, c: Optional[int] = None,
                    dtype=np.float32) -> Dict[str, np.ndarray]:
    """xs,ys~U(0,1); semi-axes a,b~U(0.02,0.22) (thin: U(0.002,0.3)); theta~U(0,pi);
    covs = R diag(a^2,b^2) R^T (rotation_matrix of utils.py:273-276); sizes = (U>0.1);
    features~N(0,1) [n, m+1, c]."""
    rng = np.random.default_rng(seed)
    xs = rng.random((n, m)); ys = rng.random((n, m))
    lo, hi = (0.002, 0.3) if thin else (0.02, 0.22)
    a = lo + (hi - lo) * rng.random((n, m)); b = lo + (hi - lo) * rng.random((n, m))
    th = np.pi * rng.random((n, m))
    cs, sn = np.cos(th), np.sin(th)
    rot = np.stack([cs, sn, -sn, cs], -1).reshape(n, m, 2, 2)
    diag = np.zeros((n, m, 2, 2)); diag[..., 0, 0] = a * a; diag[..., 1, 1] = b * b
    covs = rot @ diag @ np.swapaxes(rot, -1, -2)
    sizes = (rng.random((n, m)) > 0.1).astype(np.float32)
    out = {"xs": xs.astype(dtype), "ys": ys.astype(dtype), "covs": covs.astype(dtype), "sizes": sizes}
    if c is not None:
        out["features"] = rng.standard_normal((n, m + 1, c)).astype(dtype)
    return out
