import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from blobctrl_b200 import ops
n, k = 2, 33
levels = [(64, 64, 320), (32, 32, 640), (16, 16, 1280), (8, 8, 1280)]
g = torch.Generator().manual_seed(n * 131 + k)
scs, fts = [], []
for (h, w, c) in levels:
    sc = torch.rand(n, k, h, w, generator=g)
    scs.append((sc / sc.sum(1, keepdim=True)).cuda())
    fts.append(torch.randn(n, k, c, generator=g).cuda())
outs = ops.feature_splat_levels(scs, fts)
fused = ops.feature_splat_levels(scs, fts, engine="tensor")
for (h, w, c), sc, ft, a, b in zip(levels, scs, fts, fused, outs):
    ref = torch.einsum("nkhw,nkc->nchw", sc.double(), ft.double())
    sc_ = ref.abs().max().item()
    print((h, w, c), "fused vs per-level max diff", (a - b).abs().max().item(), " fused err/scale", ((a.double() - ref).abs().max() / sc_).item(),
          " per-level err/scale", ((b.double() - ref).abs().max() / sc_).item(), " nan:", torch.isnan(a).any().item(), torch.isnan(b).any().item())
    bad = ((a - b).abs() > 0).nonzero()
    if len(bad):
        print("   first diffs at", bad[:3].tolist(), "count", len(bad), "of", a.numel(), " chunks:", sorted(set((bad[:, 1] // 320).tolist())), "images", sorted(set(bad[:, 0].tolist())))
