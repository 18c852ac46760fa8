"""Stage 3 at small K (the pipeline's K = 1 splat, few-blob scenes): FMA engine vs tensor engine."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blobctrl_b200 import ops
def timed(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for dt in (torch.float16, torch.float32):
    for (n, k, s, c) in ((16, 1, 64, 1024), (16, 2, 64, 1024), (16, 4, 64, 320), (16, 8, 64, 320), (64, 3, 64, 320), (2, 1, 64, 1024)):
        sc = torch.rand(n, k, s, s, device="cuda").to(dt); ft = torch.randn(n, k, c, device="cuda").to(dt)
        mb = n * c * s * s * sc.element_size() / 1e6
        row = [f"{dt} N={n} K={k} {s}x{s} C={c} ({mb:.0f} MB out):"]
        for eng in ("fma", "tensor"):
            t = timed(lambda: ops.feature_splat(sc, ft, engine=eng))
            row.append(f"{eng} {t:.1f} us = {mb / t * 1e3 / 1e3:.0f} GB/s")
        print(" ".join(row))
