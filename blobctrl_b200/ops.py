"""Tensor-level operators over the C ABI: allocate outputs with torch, pass raw pointers down.

One function per C entry point (include/blobsplat.h).  The reference-signature layer
(``blobctrl_b200.utils.utils``) is built from these.  CUDA tensors only — no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _capi as C

_SELECT = {"all": C.SELECT_ALL, "fg": C.SELECT_FG, "bg": C.SELECT_BG}
_MODE = {"auto": C.COMPOSITE_AUTO, "lane_pixel": C.COMPOSITE_LANE_PIXEL, "warp_scan": C.COMPOSITE_WARP_SCAN}
_ENGINE = {"auto": C.ENGINE_AUTO, "fma": C.ENGINE_FMA, "tensor": C.ENGINE_TENSOR, "tma": C.ENGINE_TMA}

_HALF = (torch.bfloat16, torch.float16)


def canonical_blobs(xs, ys, covs, sizes) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, int, int]:
    """Normalise the reference's accepted input shapes to dense [N,M] / [N,M,2,2] tensors.

    utils.py:99-103 documents xs, ys [N,M], covs [N,M,2,2]; the reference's scripts pass xs, ys of
    shape [1] with covs [1,1,2,2] (blobctrl_inference.py:101-109), which broadcasting treats as
    [M] shared across the batch.  sizes may be [N,M] or [N,M,1] (utils.py:165-166).
    Parameter dtype follows covs: float64 stays float64, everything else computes in float32.
    """
    C.require_cuda(covs, "covs")
    if covs.ndim != 4 or covs.shape[-2:] != (2, 2):
        raise RuntimeError(f"covs must be [N, M, 2, 2], got {tuple(covs.shape)}")
    n, m = covs.shape[0], covs.shape[1]
    pdt = torch.float64 if covs.dtype == torch.float64 else torch.float32
    dev = covs.device

    def centre(t, name):
        if torch.is_tensor(t) and t.shape == (n, m) and t.dtype == pdt and t.device == dev and t.is_contiguous():
            return t                                        # already canonical: no torch ops on the latency path
        t = torch.as_tensor(t, device=dev)
        if t.ndim == 0:
            t = t.reshape(1)
        if t.ndim == 1:           # [M] (or [1]): shared across images
            t = t.unsqueeze(0)
        if t.ndim != 2:
            raise RuntimeError(f"{name} must be [N, M], got {tuple(t.shape)}")
        return t.to(pdt).expand(n, m).contiguous()

    if (torch.is_tensor(sizes) and sizes.shape == (n, m) and sizes.dtype == torch.float32 and sizes.device == dev
            and sizes.is_contiguous() and covs.dtype == pdt and covs.is_contiguous()):
        return centre(xs, "xs"), centre(ys, "ys"), covs, sizes, n, m
    sizes = torch.as_tensor(sizes, device=dev)
    if sizes.ndim == 3:
        sizes = sizes.squeeze(-1)
    if sizes.ndim == 1:
        sizes = sizes.unsqueeze(0)
    sizes = sizes.to(torch.float32).expand(n, m).contiguous()
    return centre(xs, "xs"), centre(ys, "ys"), covs.to(pdt).contiguous(), sizes, n, m


def render_scores(xs, ys, covs, sizes, height: int, width: int, select: str = "all", want_raw: bool = False,
                  want_composed: bool = True, out_dtype: Optional[torch.dtype] = None,
                  composite_mode: str = "auto") -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Stages 1+2 (blobsplat_scores).  Returns (composed [N,Ksel,H,W] | None, raw [N,K,H,W] | None)."""
    xs, ys, covs_c, sizes, n, m = canonical_blobs(xs, ys, covs, sizes)
    if out_dtype is None:
        out_dtype = covs.dtype if covs.dtype in (torch.float32, torch.float64) + _HALF else torch.float32
    if n > 65535:                                           # grid.y limit of the stand-alone kernel: split the batch
        parts = [render_scores(xs[i:i + 65535], ys[i:i + 65535], covs_c[i:i + 65535], sizes[i:i + 65535], height, width,
                               select, want_raw, want_composed, out_dtype, composite_mode) for i in range(0, n, 65535)]
        cat = lambda j: torch.cat([p[j] for p in parts]) if parts[0][j] is not None else None
        return cat(0), cat(1)
    ksel = {"all": m + 1, "fg": m, "bg": 1}[select]
    dev = covs_c.device
    composed = C.new_output((n, ksel, height, width), out_dtype, dev) if want_composed else None
    raw = C.new_output((n, m + 1, height, width), out_dtype, dev) if want_raw else None
    oc = C.dtype_code(out_dtype)
    C.check(C.lib().blobsplat_scores(C.ptr(xs), C.ptr(ys), C.ptr(covs_c), C.ptr(sizes), C.dtype_code(covs_c.dtype),
                                     n, m, height, width, _SELECT[select], C.ptr(composed), oc, C.ptr(raw), oc,
                                     _MODE[composite_mode], C.dev_of(covs_c), C.stream_of(covs_c)))
    return composed, raw


def composite(scores_nkhw: torch.Tensor) -> torch.Tensor:
    """Stage 2 alone on planar raw scores [N,K,H,W] (blobsplat_composite)."""
    C.require_cuda(scores_nkhw, "scores")
    s = scores_nkhw.contiguous()
    n, k, h, w = s.shape
    out = C.new_output(s.shape, s.dtype, s.device)
    C.check(C.lib().blobsplat_composite(C.ptr(s), C.ptr(out), n, k, h, w, C.dtype_code(s.dtype), C.dev_of(s),
                                        C.stream_of(s)))
    return out


def resize_bilinear(img: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """[N,C,H,W] -> [N,C,out_h,out_w], align_corners=False (blobsplat_resize_bilinear)."""
    C.require_cuda(img, "img")
    x = img.contiguous()
    n, c, h, w = x.shape
    out = C.new_output((n, c, out_h, out_w), x.dtype, x.device)
    C.check(C.lib().blobsplat_resize_bilinear(C.ptr(x), C.ptr(out), n * c, h, w, out_h, out_w, C.dtype_code(x.dtype),
                                              C.dev_of(x), C.stream_of(x)))
    return out


def halving_pyramid(img: torch.Tensor, cutoff: int) -> Dict[int, torch.Tensor]:
    """utils.py:280-294 semantics; even square levels go through the one-launch pyramid kernel
    (up to 3 levels per launch), anything else through the general bilinear resize."""
    C.require_cuda(img, "img")
    levels = [img]
    cur = img
    while cur.shape[-1] > cutoff:
        n, c, h, w = cur.shape
        # how many exact halvings can one launch do from here?
        todo, s = 0, w
        while h == w and s > cutoff and s % 2 == 0 and todo < 3:
            s //= 2
            todo += 1
        if todo:
            src = cur.contiguous()
            outs = [C.new_output((n, c, w >> (l + 1), w >> (l + 1)), cur.dtype, cur.device)
                    for l in range(todo)]
            arr = (ctypes.c_void_p * todo)(*[o.data_ptr() for o in outs])
            C.check(C.lib().blobsplat_pyramid(C.ptr(src), arr, todo, n * c, w, C.dtype_code(cur.dtype),
                                              C.dev_of(cur), C.stream_of(cur)))
            levels.extend(outs)
            cur = outs[-1]
        else:
            cur = resize_bilinear(cur, w // 2, w // 2)
            levels.append(cur)
    return {int(t.size(-1)): t for t in levels}


def _linear_pixel_strides(s: torch.Tensor):
    """(stride_n, stride_k, stride_p) of a logical [N,K,H,W] view whose pixel index p = y*W + x is
    linear in memory, else None (the caller then materialises a contiguous copy)."""
    _, _, h, w = s.shape
    sn, sk, sh, sw = s.stride()
    if w == 1:
        sp = sh if h > 1 else 1
    elif h == 1 or sh == w * sw:
        sp = sw
    else:
        return None
    if sp < 1 or sk < 0 or sn < 0:
        return None
    return sn, sk, sp


def feature_splat(scores: torch.Tensor, features: torch.Tensor, channels_last: bool = False,
                  engine: str = "auto") -> torch.Tensor:
    """Stage 3 (blobsplat_feature_splat): scores [N,K,H,W] (or [N,H,W,K]) x features [N,K,C] -> [N,C,H,W].
    Strided score views are consumed in place when their pixel index is linear (no copy)."""
    C.require_cuda(scores, "scores")
    if scores.ndim != 4 or features.ndim != 3:
        raise RuntimeError(f"scores must be 4-D and features 3-D, got {tuple(scores.shape)} / {tuple(features.shape)}")
    s = scores.permute(0, 3, 1, 2) if channels_last else scores        # logical [N,K,H,W]
    n, k, h, w = s.shape
    if features.shape[0] != n or features.shape[1] != k:
        raise RuntimeError(f"einsum(): operands do not broadcast: scores {tuple(s.shape)} (N,K,H,W) vs features "
                           f"{tuple(features.shape)} (N,K,C)")
    f = features.to(dtype=s.dtype, device=s.device).contiguous()        # utils.py:69
    strides = _linear_pixel_strides(s)
    if strides is None:
        s = s.contiguous()
        strides = _linear_pixel_strides(s)
    sn, sk, sp = strides
    c = f.shape[2]
    out = C.new_output((n, c, h, w), s.dtype, s.device)
    C.check(C.lib().blobsplat_feature_splat(C.ptr(s), sn, sk, sp, C.ptr(f), C.ptr(out), n, k, c, h, w,
                                            C.dtype_code(s.dtype), _ENGINE[engine], C.dev_of(s), C.stream_of(s)))
    return out


def feature_splat_levels(scores: Sequence[torch.Tensor], features: Sequence[torch.Tensor],
                         engine: str = "auto") -> List[torch.Tensor]:
    """Stage 3 for a whole pyramid (blobsplat_feature_splat_levels): scores[i] [N,K,H_i,W_i] x features[i] [N,K,C_i]
    -> [N,C_i,H_i,W_i] for every level, as ONE tcgen05 launch when the levels share an operand tiling (BlobNet's
    320/640/1280-channel levels do), level by level otherwise.  Same dtype / N / K on every level."""
    if len(scores) != len(features):
        raise RuntimeError("one feature tensor per score map")
    if not scores:
        return []
    C.require_cuda(scores[0], "scores")
    n, k = scores[0].shape[:2]
    dt, dev = scores[0].dtype, scores[0].device
    ss, fs, outs, strides, shapes = [], [], [], [], []
    for s, f in zip(scores, features):
        if s.ndim != 4 or f.ndim != 3 or s.shape[0] != n or s.shape[1] != k or s.dtype != dt or s.device != dev:
            raise RuntimeError("levels must be [N,K,H,W] maps of one batch, dtype and device")
        if f.shape[0] != n or f.shape[1] != k:
            raise RuntimeError(f"einsum(): operands do not broadcast: scores {tuple(s.shape)} (N,K,H,W) vs features "
                               f"{tuple(f.shape)} (N,K,C)")
        st = _linear_pixel_strides(s)
        if st is None:
            s = s.contiguous()
            st = _linear_pixel_strides(s)
        f = f.to(dtype=dt, device=dev).contiguous()
        ss.append(s); fs.append(f); strides.append(st); shapes.append((f.shape[2], s.shape[2], s.shape[3]))
        outs.append(C.new_output((n, f.shape[2], s.shape[2], s.shape[3]), dt, dev))
    L = len(ss)
    ptrs = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])
    i64 = lambda vals: (ctypes.c_int64 * L)(*vals)
    i32 = lambda vals: (ctypes.c_int * L)(*vals)
    C.check(C.lib().blobsplat_feature_splat_levels(
        L, ptrs(ss), i64([st[0] for st in strides]), i64([st[1] for st in strides]), i64([st[2] for st in strides]),
        ptrs(fs), ptrs(outs), n, k, i32([sh[0] for sh in shapes]), i32([sh[1] for sh in shapes]),
        i32([sh[2] for sh in shapes]), C.dtype_code(dt), _ENGINE[engine], C.dev_of(ss[0]), C.stream_of(ss[0])))
    return outs


def render_fused(xs, ys, covs, sizes, features: torch.Tensor, height: int, width: int,
                 out_dtype: Optional[torch.dtype] = None, want_composed: bool = True):
    """Stages 1+2+3 in one launch (blobsplat_render, tcgen05).  Returns (composed | None, grid).
    Raises BlobSplatUnsupported when the shape is outside the tensor-core kernel's envelope."""
    xs, ys, covs_c, sizes, n, m = canonical_blobs(xs, ys, covs, sizes)
    if covs_c.dtype != torch.float32:
        raise C.BlobSplatUnsupported("blobsplat: unsupported: fused render takes float32 blob parameters")
    if out_dtype is None:
        out_dtype = features.dtype
    f = features
    if f.dtype != out_dtype or f.device != covs_c.device or not f.is_contiguous():
        f = features.to(device=covs_c.device, dtype=out_dtype).contiguous()     # utils.py:69
    if f.ndim != 3 or f.shape[0] != n or f.shape[1] != m + 1:
        raise RuntimeError(f"features must be [N, M+1, C] = [{n}, {m + 1}, C], got {tuple(f.shape)}")
    c = f.shape[2]
    dev = covs_c.device
    composed = C.new_output((n, m + 1, height, width), out_dtype, dev) if want_composed else None
    grid = C.new_output((n, c, height, width), out_dtype, dev)
    C.check(C.lib().blobsplat_render(C.ptr(xs), C.ptr(ys), C.ptr(covs_c), C.ptr(sizes), C.ptr(f),
                                     C.dtype_code(f.dtype), n, m, height, width, c, C.ptr(composed), C.ptr(grid),
                                     C.dtype_code(out_dtype), C.dev_of(covs_c), C.stream_of(covs_c)))
    return composed, grid


SMALL_RENDER_MAX_K = 17            # AUTO takes the latency kernel up to 17 planes (it renders up to 33: the per-pixel weights live
                                   # in registers); with more blobs every channel tile's repeat of stages 1+2 outweighs the start-up it saves
SMALL_RENDER_MAX_WORK = 3 << 23    # and up to 25 M multiply-adds (P * K * C) of the ONE image.  Graph replays on a B200, latency vs
                                   # tensor kernel: cfg2 (22 M) 8.2 vs 12.1 us, K = 2 x C = 1024 8.2 vs 14.3, M = 8 at 128 x 128 x 64
                                   # 6.2 vs 10.3, M = 4 at 256 x 256 x 64 10.3 vs 14.3; M = 32 x C = 160 12.3 vs 10.3 (profiles/ab_small_r2.txt)


def small_render_applies(n: int, m: int, height: int, width: int, c: int) -> bool:
    """The one-image latency rule: a single float32 image whose stage 3 is a few microseconds of FP32 work (BASELINE config 2
    is 22 M multiply-adds).  Batches never take it, so an image's bits do not depend on the size of the batch it is in —
    a lone image is simply rendered with full-fp32 products instead of the split-precision tensor path (both within 1e-6)."""
    return n == 1 and m + 1 <= SMALL_RENDER_MAX_K and height * width * (m + 1) * c <= SMALL_RENDER_MAX_WORK


def render_small(xs, ys, covs, sizes, features: torch.Tensor, height: int, width: int, want_composed: bool = True):
    """Stages 1+2+3 in one CUDA-core launch (blobsplat_render_small): the latency kernel for single small float32 renders.
    Returns (composed | None, grid).  Raises BlobSplatUnsupported outside its envelope (float32, K <= 33)."""
    xs, ys, covs_c, sizes, n, m = canonical_blobs(xs, ys, covs, sizes)
    if covs_c.dtype != torch.float32 or features.dtype != torch.float32:
        raise C.BlobSplatUnsupported("blobsplat: unsupported: the latency render takes float32 parameters and features")
    f = features
    if f.device != covs_c.device or not f.is_contiguous():
        f = features.to(device=covs_c.device).contiguous()
    if f.ndim != 3 or f.shape[0] != n or f.shape[1] != m + 1:
        raise RuntimeError(f"features must be [N, M+1, C] = [{n}, {m + 1}, C], got {tuple(f.shape)}")
    c = f.shape[2]
    dev = covs_c.device
    composed = C.new_output((n, m + 1, height, width), torch.float32, dev) if want_composed else None
    grid = C.new_output((n, c, height, width), torch.float32, dev)
    C.check(C.lib().blobsplat_render_small(C.ptr(xs), C.ptr(ys), C.ptr(covs_c), C.ptr(sizes), C.ptr(f), n, m, height, width, c,
                                           C.ptr(composed), C.ptr(grid), C.dev_of(covs_c), C.stream_of(covs_c)))
    return composed, grid


def render_multiscale(xs, ys, covs, sizes, size: int, level_features: Sequence[Optional[torch.Tensor]],
                      out_dtype: torch.dtype):
    """blobsplat_render_multiscale: level l has size ``size >> l``; level_features[l] is [N, M+1, C_l] or None (maps only).
    Returns (composed maps per level, grids per level (None where no features)).  One C call for the whole pyramid."""
    xs, ys, covs_c, sizes, n, m = canonical_blobs(xs, ys, covs, sizes)
    if covs_c.dtype != torch.float32:
        raise C.BlobSplatUnsupported("blobsplat: unsupported: fused render takes float32 blob parameters")
    dev = covs_c.device
    L = len(level_features)
    feats, comps, grids, cs = [], [], [], []
    for l, f in enumerate(level_features):
        s = size >> l
        comps.append(C.new_output((n, m + 1, s, s), out_dtype, dev))
        if f is None:
            feats.append(None); grids.append(None); cs.append(0)
            continue
        f = f.to(device=dev, dtype=out_dtype).contiguous()
        if f.ndim != 3 or f.shape[0] != n or f.shape[1] != m + 1:
            raise RuntimeError(f"features must be [N, M+1, C] = [{n}, {m + 1}, C], got {tuple(f.shape)}")
        feats.append(f); cs.append(f.shape[2])
        grids.append(C.new_output((n, f.shape[2], s, s), out_dtype, dev))
    vp = lambda ts: (ctypes.c_void_p * L)(*[None if t is None else t.data_ptr() for t in ts])
    C.check(C.lib().blobsplat_render_multiscale(C.ptr(xs), C.ptr(ys), C.ptr(covs_c), C.ptr(sizes), n, m, size, L, vp(feats),
                                                (ctypes.c_int * L)(*cs), vp(comps), vp(grids), C.dtype_code(out_dtype),
                                                C.dev_of(covs_c), C.stream_of(covs_c)))
    return comps, grids


def render_fused_into(xs, ys, covs, sizes, features, height: int, width: int, composed: Optional[torch.Tensor],
                      grid: torch.Tensor) -> None:
    """blobsplat_render into caller-owned (contiguous) output buffers; inputs must already be dense float32 [N,M]
    device tensors (no canonicalisation, no allocation: usable under CUDA-graph capture and in copy pipelines)."""
    n, m = covs.shape[0], covs.shape[1]
    c = features.shape[2]
    for t in (xs, ys, covs, sizes, features, grid) + ((composed,) if composed is not None else ()):
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("render_fused_into needs contiguous CUDA tensors")
    if grid.shape != (n, c, height, width) or (composed is not None and composed.shape != (n, m + 1, height, width)):
        raise RuntimeError("output buffer shape mismatch")
    if features.dtype != grid.dtype or (composed is not None and composed.dtype != grid.dtype):
        raise RuntimeError("features, composed and grid must share a dtype")
    C.check(C.lib().blobsplat_render(C.ptr(xs), C.ptr(ys), C.ptr(covs), C.ptr(sizes), C.ptr(features),
                                     C.dtype_code(features.dtype), n, m, height, width, c, C.ptr(composed), C.ptr(grid),
                                     C.dtype_code(grid.dtype), C.dev_of(grid), C.stream_of(grid)))


def render_scores_from_ellipses(ellipses: torch.Tensor, sizes: Optional[torch.Tensor], image_size: Tuple[int, int],
                                height: int, width: int, select: str = "all", want_raw: bool = False,
                                out_dtype: torch.dtype = torch.float32):
    """Stages 1+2 straight from OpenCV ellipses (blobsplat_scores_ellipse).

    ellipses: [N, M, 5] (xc, yc, d1, d2, angle_deg) in pixels of an image of ``image_size`` = (img_h, img_w);
    sizes: [N, M] existence flags or None (all blobs exist).  Returns (composed [N,Ksel,H,W], raw | None)."""
    C.require_cuda(ellipses, "ellipses")
    if ellipses.ndim != 3 or ellipses.shape[-1] != 5:
        raise RuntimeError(f"ellipses must be [N, M, 5], got {tuple(ellipses.shape)}")
    n, m = ellipses.shape[:2]
    e = ellipses.to(torch.float32).contiguous()
    sz = torch.ones((n, m), dtype=torch.float32, device=e.device) if sizes is None else \
        torch.as_tensor(sizes, device=e.device).to(torch.float32).reshape(n, m).contiguous()
    img_h, img_w = image_size
    ksel = {"all": m + 1, "fg": m, "bg": 1}[select]
    composed = C.new_output((n, ksel, height, width), out_dtype, e.device)
    raw = C.new_output((n, m + 1, height, width), out_dtype, e.device) if want_raw else None
    oc = C.dtype_code(out_dtype)
    C.check(C.lib().blobsplat_scores_ellipse(C.ptr(e), C.ptr(sz), float(img_w), float(img_h), n, m, height, width,
                                             _SELECT[select], C.ptr(composed), oc, C.ptr(raw), oc, C.dev_of(e),
                                             C.stream_of(e)))
    return composed, raw


def render_preview(xs, ys, covs, sizes, colors: torch.Tensor, height: int, width: int, want_composed: bool = False):
    """Stages 1+2 + the C = 3 colour splat in one launch (blobsplat_preview): image [N,3,H,W] = sum_k d_k * colors[k]
    without materialising the score maps.  colors: [K, 3] or [N, K, 3] (K = M + 1 rows are read).  float32 or float64
    (the dtype of ``covs``).  Returns (image, composed | None)."""
    xs, ys, covs_c, sizes, n, m = canonical_blobs(xs, ys, covs, sizes)
    dt, dev = covs_c.dtype, covs_c.device
    col = colors.to(device=dev, dtype=dt)
    if col.ndim == 2:
        col, per_image = col[: m + 1].contiguous(), 0
    elif col.ndim == 3 and col.shape[0] == n:
        col, per_image = col[:, : m + 1].contiguous(), 1
    else:
        raise RuntimeError(f"colors must be [K, 3] or [N, K, 3], got {tuple(colors.shape)}")
    if col.shape[-2] < m + 1 or col.shape[-1] != 3:
        raise RuntimeError(f"colors must hold {m + 1} RGB rows, got {tuple(colors.shape)}")
    image = C.new_output((n, 3, height, width), dt, dev)
    composed = C.new_output((n, m + 1, height, width), dt, dev) if want_composed else None
    C.check(C.lib().blobsplat_preview(C.ptr(xs), C.ptr(ys), C.ptr(covs_c), C.ptr(sizes), C.dtype_code(dt), C.ptr(col), per_image,
                                      n, m, height, width, C.ptr(image), C.ptr(composed), C.dev_of(covs_c), C.stream_of(covs_c)))
    return image, composed


def render_preview_u8(xs, ys, covs, sizes, colors: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """The preview as the 8-bit picture the UI shows (blobsplat_preview_u8): [N, H, W, 3] uint8 =
    (image.permute(0, 2, 3, 1) * 255) truncated like numpy's astype(np.uint8) (scripts/blobctrl_app.py:645-646), straight
    from the render launch — a quarter (float32) or an eighth (float64) of the bytes to bring back to the host."""
    xs, ys, covs_c, sizes, n, m = canonical_blobs(xs, ys, covs, sizes)
    dt, dev = covs_c.dtype, covs_c.device
    col = colors.to(device=dev, dtype=dt)
    if col.ndim == 2:
        col, per_image = col[: m + 1].contiguous(), 0
    elif col.ndim == 3 and col.shape[0] == n:
        col, per_image = col[:, : m + 1].contiguous(), 1
    else:
        raise RuntimeError(f"colors must be [K, 3] or [N, K, 3], got {tuple(colors.shape)}")
    if col.shape[-2] < m + 1 or col.shape[-1] != 3:
        raise RuntimeError(f"colors must hold {m + 1} RGB rows, got {tuple(colors.shape)}")
    image = torch.empty((n, height, width, 3), dtype=torch.uint8, device=dev)
    C.check(C.lib().blobsplat_preview_u8(C.ptr(xs), C.ptr(ys), C.ptr(covs_c), C.ptr(sizes), C.dtype_code(dt), C.ptr(col), per_image,
                                         n, m, height, width, C.ptr(image), C.dev_of(covs_c), C.stream_of(covs_c)))
    return image


def conv_in_weights(weight: torch.Tensor, features: torch.Tensor, latent_channels: int = 4) -> torch.Tensor:
    """blobsplat_conv_in_weights: per-sample effective 3x3 kernels of the conditioning planes of BlobNet's conv_in.
    weight [O, lc+1+C, 3, 3], features [B, K, C] -> weff [B, O, 1+K, 12] float32 (9 taps padded to 12)."""
    C.require_cuda(weight, "weight")
    w = weight.contiguous()
    f = features.to(device=w.device, dtype=w.dtype).contiguous()
    o, cin = w.shape[:2]
    b, k, c = f.shape
    if w.shape[2:] != (3, 3) or cin != latent_channels + 1 + c:
        raise RuntimeError(f"conv_in weight {tuple(w.shape)} does not match {latent_channels} latent + 1 score + {c} feature planes")
    weff = torch.zeros((b, o, 1 + k, 12), dtype=torch.float32, device=w.device)
    C.check(C.lib().blobsplat_conv_in_weights(C.ptr(w), C.ptr(f), C.ptr(weff), b, o, cin, latent_channels, c, k,
                                              C.dtype_code(w.dtype), C.dev_of(w), C.stream_of(w)))
    return weff


def conv_in_hoisted(latents: torch.Tensor, cond: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                    weff: torch.Tensor, halves: int = 2, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """blobsplat_conv_in_hoisted: latents [B, lc, h, halves*w], cond [B, J, h, w], weight [O, Cin, 3, 3], weff [B, O, J, 12]
    -> [B, O, h, halves*w] == conv_in(construct_blobnet_input(...)) up to summation order."""
    C.require_cuda(latents, "latents")
    dt = latents.dtype
    b, lc, h, wt = latents.shape
    j, w = cond.shape[1], cond.shape[3]
    o, cin = weight.shape[:2]
    if wt != halves * w or cond.shape[0] != b or cond.shape[2] != h or tuple(weff.shape) != (b, o, j, 12):
        raise RuntimeError("conv_in_hoisted: shape mismatch between latents, cond and weff")
    for t in (latents, cond, weight, weff) + ((bias,) if bias is not None else ()):
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("conv_in_hoisted needs contiguous CUDA tensors")
    if cond.dtype != dt or weight.dtype != dt or (bias is not None and bias.dtype != dt) or weff.dtype != torch.float32:
        raise RuntimeError("conv_in_hoisted: latents, cond, weight and bias must share a dtype; weff is float32")
    if out is None:
        out = C.new_output((b, o, h, wt), dt, latents.device)
    C.check(C.lib().blobsplat_conv_in_hoisted(C.ptr(latents), C.ptr(cond), C.ptr(weight), C.ptr(bias), C.ptr(weff), C.ptr(out), b, o,
                                              cin, lc, j, h, w, halves, C.dtype_code(dt), C.dev_of(latents), C.stream_of(latents)))
    return out
