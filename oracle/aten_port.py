"""TEST INFRASTRUCTURE ONLY — PyTorch-CPU port of the reference renderer, for timing.

``/root/reference`` cannot travel to the GPU box, so the "reference's PyTorch CPU renderer timed
on the box's own host cores" (BASELINE.json north_star) is this port: the same ATen operator
sequence the reference issues (batched ``linalg.solve``, ``sigmoid``, reversed ``cumprod``,
``interpolate(bilinear)``, ``einsum``), restated stage by stage from
``/root/reference/blobctrl/utils/utils.py`` (cited as ``utils.py:<lines>``).  Because it runs the
same ATen kernels it is bit-identical to the reference on CPU; ``tests/test_oracle_golden.py``
checks that against the golden fixtures, and ``tests/golden/make_golden.py`` against the live
reference.  Used only by ``bench.py`` (``cpu_baseline``, ``--impl reference``) and tests.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as TF


def pixel_grid(height: int, width: int, device=None) -> torch.Tensor:
    """utils.py:123-125 / :139-141 — int64 [2, H*W]: row 0 = p % W, row 1 = p // W."""
    gx = torch.arange(width).repeat(height)
    gy = torch.arange(height).repeat_interleave(width)
    return torch.stack((gx, gy), 0).to(device)


def sq_mahalanobis(xs, ys, covs, height: int, width: int, square: bool) -> torch.Tensor:
    """utils.py:138-143 (square: scale both axes by S) / :147-156 (tuple: per-axis)."""
    grid = pixel_grid(height, width, xs.device)
    if square:
        centre = torch.stack((xs, ys), -1).mul(width)
        delta = (grid[None, None] - centre[..., None]).div(width)
    else:
        centre = torch.stack((xs.mul(width), ys.mul(height)), -1)
        delta = grid[None, None] - centre[..., None]
        delta[:, :, 0, :] /= width
        delta[:, :, 1, :] /= height
    return (delta * torch.linalg.solve(covs, delta)).sum(2)          # [N, M, P]


def opacity(q_nhwm: torch.Tensor, sizes: torch.Tensor) -> torch.Tensor:
    """utils.py:162-172: s = min(1, 2*sigmoid(-q)); sizes < 0.5 -> 1e-6."""
    s = q_nhwm.div(-1).sigmoid().mul(2).clamp_(max=1)
    if sizes.ndim == 3:
        sizes = sizes.squeeze(-1)
    gone = (sizes < 0.5)[:, None, None, :].expand(-1, s.shape[1], s.shape[2], -1)
    return torch.where(gone, torch.tensor(1e-6, device=s.device), s)


def alpha_composite(s: torch.Tensor) -> torch.Tensor:
    """utils.py:179-181: flip -> cumprod(1 - s) -> flip -> roll(-1) -> * s; last channel raw."""
    order = list(range(s.size(-1) - 1, -1, -1))
    d = (1 - s[..., order]).cumprod(-1)[..., order].roll(-1, -1) * s
    d[..., -1] = s[..., -1]
    return d


def halve_pyramid(img: torch.Tensor, cutoff: int) -> Dict[int, torch.Tensor]:
    """utils.py:280-294."""
    levels = [img]
    while img.shape[-1] > cutoff:
        img = TF.interpolate(img, img.shape[-1] // 2, mode="bilinear", align_corners=False)
        levels.append(img)
    return {t.size(-1): t for t in levels}


def feature_splat(scores: torch.Tensor, features: torch.Tensor, size, channels_last: bool = True):
    """utils.py:57-77."""
    features = features.to(dtype=scores.dtype, device=scores.device)
    if size and not (scores.shape[2] == size):
        if channels_last:
            scores = scores.permute(0, 3, 1, 2)
        scores = TF.interpolate(scores, size, mode="bilinear", align_corners=False)
        spec = "nmhw,nmc->nchw"
    else:
        spec = "nhwm,nmc->nchw" if channels_last else "nmhw,nmc->nchw"
    return torch.einsum(spec, scores, features).contiguous()


def render(xs, ys, covs, sizes, score_size=None, interp_size=None, features=None, viz_size=None,
           is_viz=False, ret_layout=True, viz_score_fn=None, return_d_score=False, only_vis=False,
           only_splatting_fg=False, only_splatting_bg=False, viz_colors=None):
    """utils.py:80-241, same kwargs and return shapes."""
    if viz_size is not None and not isinstance(viz_size, int):
        h, w = viz_size
        q = sq_mahalanobis(xs, ys, covs, h, w, square=False).view(1, 1, h, w).permute(0, 2, 3, 1).contiguous()
    elif isinstance(score_size, int):
        h = w = score_size
        q = sq_mahalanobis(xs, ys, covs, h, w, square=True)
        q = q.view(q.shape[0], q.shape[1], h, w).permute(0, 2, 3, 1)
    else:
        h, w = score_size
        q = sq_mahalanobis(xs, ys, covs, h, w, square=False).view(1, 1, h, w).permute(0, 2, 3, 1).contiguous()
    s = opacity(q, sizes)
    s = torch.cat((torch.ones_like(s[..., :1]), s), -1)               # utils.py:175-176
    d = alpha_composite(s)
    if only_splatting_bg:
        d = d[..., 0].unsqueeze(-1)
    elif only_splatting_fg:
        d = d[..., 1:]
    if return_d_score:
        return d.permute(0, 3, 1, 2)
    out = {}
    if is_viz:
        sv = alpha_composite(viz_score_fn(s)) if viz_score_fn is not None else d
        k = xs.shape[-1] + 1
        vc = viz_colors.to(sv.device)
        vc = vc[:k][None].repeat_interleave(len(sv), 0) if vc.ndim == 2 else vc[:, :k]
        out["feature_img"] = feature_splat(sv, vc, viz_size)
    if only_vis:
        return out
    out["scores_pyramid"] = halve_pyramid(d.permute(0, 3, 1, 2), interp_size)
    out["feature_grid"] = feature_splat(out["scores_pyramid"][interp_size], features, interp_size, False)
    out.update({"feature_img": None, "entropy_img": None})
    if ret_layout:
        out.update({"xs": xs, "ys": ys, "covs": covs, "raw_scores": s, "sizes": sizes,
                    "composed_scores": d, "features": features})
    return out


def multiscale(xs, ys, covs, sizes, score_size: int, level_features: Dict[int, torch.Tensor]):
    """BASELINE config 3: render at score_size, halve down to the smallest
    level, and splat per-level features [N, M+1, C_S] at every level S."""
    d = render(xs, ys, covs, sizes, score_size=score_size, return_d_score=True)
    pyr = halve_pyramid(d, min(level_features))
    grids = {s: feature_splat(pyr[s], f, s, channels_last=False) for s, f in level_features.items()}
    return {"scores_pyramid": pyr, "feature_grids": grids}
