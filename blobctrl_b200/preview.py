"""CUDA-graph renderers for the latency-bound uses of the path (BASELINE configs 1-2).

The reference's UI re-renders the blob preview on every mouse event (scripts/blobctrl_app.py:637-650, called from
:908, :1103, :1151, :1217): one image, one blob, 512x512, and the CLI renders one 64x64 score map per edit
(scripts/blobctrl_inference.py:112-117).  At those sizes the kernels take a few microseconds and the cost is host
launch overhead, so the whole render (stages 1+2, and stage 3 with the palette or the features) is captured once in a
CUDA graph over static buffers; a call copies the ~28 bytes per blob of parameters into place and replays it.
The C ABI allocates nothing and never synchronises, which is what makes the capture legal.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _capi as C
from . import ops
from .utils.utils import BLOB_VIS_COLORS


class GraphedBlobRenderer:
    """Fixed-shape renderer: N images x M blobs at (H, W); optional stage 3 with ``features`` [N, M+1, C] given per
    call, or with a fixed colour table (``viz_colors``: the UI preview).  float32 parameters and maps."""

    def __init__(self, n: int, m: int, size: Tuple[int, int], channels: Optional[int] = None,
                 viz_colors: Optional[torch.Tensor] = None, device="cuda", picture: bool = False):
        self.n, self.m, (self.h, self.w) = n, m, size
        # picture: the preview leaves the launch as the 8-bit [N, H, W, 3] image the UI shows (blobsplat_preview_u8) and the
        # graph ends with its copy into pinned host memory — one replay is parameters in, picture out (render_picture)
        self.picture = bool(picture) and viz_colors is not None
        self.pic = None
        self.pic_host = torch.empty((n, size[0], size[1], 3), dtype=torch.uint8).pin_memory() if self.picture else None
        dev = torch.device(device)
        f32 = dict(dtype=torch.float32, device=dev)
        # one flat parameter block (xs | ys | covs | sizes) so a call is ONE host->device copy + one graph replay
        nm = n * m
        self.params = torch.zeros(7 * nm, **f32)
        self.xs = self.params[:nm].view(n, m); self.ys = self.params[nm:2 * nm].view(n, m)
        self.covs = self.params[2 * nm:6 * nm].view(n, m, 2, 2); self.sizes = self.params[6 * nm:].view(n, m)
        # pinned staging is double-buffered: block i is rewritten only after the H2D copy that last read it has run
        # (an event per block), so back-to-back calls without a sync never render call i with call i+1's parameters
        self._hosts = [torch.zeros(7 * nm, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._hs = [h.numpy() for h in self._hosts]
        self._copied = [torch.cuda.Event(), torch.cuda.Event()]
        self._turn = 0
        for h in self._hs:
            h[2 * nm:6 * nm] = [1, 0, 0, 1] * nm; h[6 * nm:] = 1
        self.params.copy_(self._hosts[0])
        self.feats = None
        self.colors = None
        if viz_colors is not None:
            self.colors = viz_colors[: m + 1].to(**f32).contiguous()                          # utils.py:251-253
        elif channels:
            self.feats = torch.zeros((n, m + 1, channels), **f32)
        self.takes_features = viz_colors is None and bool(channels)
        self.stream = torch.cuda.Stream(device=dev)
        self._run()                                      # warm-up outside capture (loads the kernels)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._run()

    def _run(self):
        with torch.cuda.stream(self.stream):
            c = self.feats.shape[-1] if self.feats is not None else 0
            self.grid = None
            if self.picture:
                self.scores = None
                self.pic = ops.render_preview_u8(self.xs, self.ys, self.covs, self.sizes, self.colors, self.h, self.w)
                self.pic_host.copy_(self.pic, non_blocking=True)
                return
            if self.colors is not None:      # the UI preview: stages 1+2 + colour splat in one launch (blobsplat_preview)
                self.grid, self.scores = ops.render_preview(self.xs, self.ys, self.covs, self.sizes, self.colors, self.h, self.w,
                                                            want_composed=True)
                return
            if self.feats is not None and not C.f32_exact() and ops.small_render_applies(self.n, self.m, self.h, self.w, c):
                # the same rule as splat_features: a single small image takes the CUDA-core latency kernel
                self.scores, self.grid = ops.render_small(self.xs, self.ys, self.covs, self.sizes, self.feats, self.h, self.w)
                return
            if self.feats is not None and c >= 64:
                try:
                    self.scores, self.grid = ops.render_fused(self.xs, self.ys, self.covs, self.sizes, self.feats, self.h, self.w)
                    return
                except C.BlobSplatUnsupported:   # outside the tensor-core envelope only; CUDA failures propagate
                    pass
            self.scores, _ = ops.render_scores(self.xs, self.ys, self.covs, self.sizes, self.h, self.w)
            if self.feats is not None:
                self.grid = ops.feature_splat(self.scores, self.feats)

    @torch.no_grad()
    def __call__(self, xs, ys, covs, sizes=None, features=None):
        """Parameters: host arrays / tensors of any float dtype (the reference's callers build them with numpy on the
        host, blobctrl_inference.py:101-109); device tensors are accepted but cost a sync.  Returns (composed [N,M+1,H,W], grid | None) —
        views of the renderer's static output buffers: the next call overwrites them in stream order, so consume (or clone)
        them on the current stream before calling again."""
        import numpy as np
        nm = self.n * self.m
        to_np = lambda t: t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
        i = self._turn
        self._turn ^= 1
        self._copied[i].synchronize()               # the copy that last read this pinned block has completed
        h = self._hs[i]
        h[:nm] = to_np(xs).reshape(-1); h[nm:2 * nm] = to_np(ys).reshape(-1)
        h[2 * nm:6 * nm] = to_np(covs).reshape(-1)
        h[6 * nm:] = 1 if sizes is None else to_np(sizes).reshape(-1)
        cur = torch.cuda.current_stream(self.params.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.params.copy_(self._hosts[i], non_blocking=True)
            self._copied[i].record(self.stream)
            if features is not None and self.takes_features:
                self.feats.copy_(features, non_blocking=True)
            self.graph.replay()
        cur.wait_stream(self.stream)
        return (None, self.pic) if self.picture else (self.scores, self.grid)

    def render_picture(self, xs, ys, covs, sizes=None):
        """The UI's whole preview step (scripts/blobctrl_app.py:637-648 up to Image.fromarray), synchronous: host parameters
        in, host uint8 picture [H, W, 3] of image 0 out (a view of the renderer's pinned buffer, rewritten by the next call)."""
        if not self.picture:
            raise RuntimeError("construct the renderer with picture=True (and viz_colors)")
        self(xs, ys, covs, sizes)
        self.stream.synchronize()
        return self.pic_host[0].numpy()


def preview_renderer(viz_size=(512, 512), device="cuda", picture: bool = False) -> GraphedBlobRenderer:
    """The UI preview of scripts/blobctrl_app.py:637-646 (one image, one blob, BLOB_VIS_COLORS) as a graph:
    ``scores, img = r(xs, ys, covs)`` -> img [1, 3, H, W] in [0, 1]; with ``picture=True``
    ``r.render_picture(xs, ys, covs)`` -> the host uint8 [H, W, 3] array ``Image.fromarray`` takes (:647-648)."""
    return GraphedBlobRenderer(1, 1, viz_size, viz_colors=BLOB_VIS_COLORS, device=device, picture=picture)
