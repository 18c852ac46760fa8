"""Latency of the small launches (cfg2: 1 image, 16 blobs, 64x64, C = 320) — eager and as a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blobctrl_b200 as B
from blobctrl_b200 import ops
from bench import synthetic
dev = "cuda"
def timed(fn, reps=200, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for n, m, c, dt in ((1, 16, 320, torch.float32), (1, 16, 320, torch.bfloat16), (2, 16, 320, torch.float32), (4, 64, 320, torch.float32),
                    (1, 64, 1280, torch.bfloat16), (8, 32, 320, torch.bfloat16)):
    hb, hf = synthetic(n, m, c, seed=0)
    b2 = {k: v.to(dev) for k, v in hb.items()}; f2 = hf.to(dev).to(dt)
    kw = dict(features=f2, score_size=64, interp_size=64, ret_layout=False, out_dtype=dt)
    e = timed(lambda: B.splat_features(**b2, **kw), reps=50)
    g = timed(lambda: B.splat_features(**b2, **kw, cuda_graph=True))
    xs, ys, covs, sizes, _, _ = ops.canonical_blobs(**b2)
    comp = torch.empty((n, m + 1, 64, 64), dtype=dt, device=dev); grid = torch.empty((n, c, 64, 64), dtype=dt, device=dev)
    fn = lambda: ops.render_fused_into(xs, ys, covs, sizes, f2, 64, 64, comp, grid)
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    k = timed(gr.replay)
    print(f"N={n} M={m} C={c} {str(dt)[6:]:9s}: splat_features eager {e:6.1f} us  cuda_graph=True {g:6.1f} us   bare render graph {k:6.1f} us", flush=True)
