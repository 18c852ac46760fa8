"""A/B the fused render's compute-warp count (BLOBSPLAT_TC_HALVES) in one process, interleaved."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic
from blobctrl_b200 import ops

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

blobs, feats = synthetic(1024, 64, 320, seed=0)
b = {k: v.cuda() for k, v in blobs.items()}
for dtype in (torch.float32, torch.bfloat16):
    f = feats.cuda().to(dtype)
    res = {"1": [], "2": [], "4": []}
    for rnd in range(4):
        for h in ("1", "2", "4"):
            os.environ["BLOBSPLAT_TC_HALVES"] = h
            res[h].append(t(lambda: ops.render_fused(**b, features=f, height=64, width=64, out_dtype=dtype)))
    print(dtype, {h: [round(x, 4) for x in v] for h, v in res.items()})
