"""Per-level stage-3 timing (cfg3 shapes, bf16): tensor vs FMA engine, CUDA-graph replay of 20 calls each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blobctrl_b200 import ops
torch.manual_seed(0)
def timed(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(reps): f()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for n in (64, 8, 1):
    for s, c in ((64, 320), (32, 640), (16, 1280), (8, 1280)):
        sc = torch.rand(n, 33, s, s, device="cuda"); sc = (sc / sc.sum(1, keepdim=True)).to(torch.bfloat16)
        ft = torch.randn(n, 33, c, device="cuda").to(torch.bfloat16)
        mb = (n * c * s * s * 2 + sc.numel() * 2 + ft.numel() * 2) / 1e6
        row = [f"N={n} {s}x{s} C={c} ({mb:.0f} MB):"]
        for eng in ("tensor", "fma"):
            try:
                t = timed(lambda: ops.feature_splat(sc, ft, engine=eng))
                row.append(f"{eng} {t:.1f} us ({mb / t / 1e3 * 1e3:.0f} GB/s)")
            except Exception as e:
                row.append(f"{eng} failed: {str(e)[:40]}")
        print(" ".join(row))

# the three lower cfg3 levels: per-level launches (engine auto) vs ONE launch (engine tensor)
n = 64
scs, fts = [], []
for s, c in ((32, 640), (16, 1280), (8, 1280)):
    sc = torch.rand(n, 33, s, s, device="cuda"); scs.append((sc / sc.sum(1, keepdim=True)).to(torch.bfloat16))
    fts.append(torch.randn(n, 33, c, device="cuda").to(torch.bfloat16))
for eng in ("auto", "tensor"):
    print(f"cfg3 levels 32/16/8 engine={eng}: {timed(lambda: ops.feature_splat_levels(scs, fts, engine=eng)):.1f} us")
