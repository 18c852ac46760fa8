"""A/B of the two float32 split-precision forms of the fused render (3xTF32 vs 2xFP16) at the bench shape: burst (20
launches) and sustained (2 s back to back), alternating, plus accuracy of both against a float64 contraction."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic
from blobctrl_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
blobs, feats = synthetic(n, 64, 320, seed=0)
b = {k: v.cuda() for k, v in blobs.items()}
f = feats.cuda()
comp = torch.empty((n, 65, 64, 64), device="cuda"); grid = torch.empty((n, 320, 64, 64), device="cuda")
call = lambda: ops.render_fused_into(b["xs"], b["ys"], b["covs"], b["sizes"], f, 64, 64, comp, grid)


def burst(reps=20, warm=5):
    for _ in range(warm):
        call()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        call()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / reps


def sustained(seconds=2.0):
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, calls = time.perf_counter(), 0
    a.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            call()
        calls += 50
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / calls


def set_split(s):
    if s == "f16":
        os.environ.pop("BLOBSPLAT_F32_SPLIT", None)
    else:
        os.environ["BLOBSPLAT_F32_SPLIT"] = s


bytes_ = n * (28 * 64 + 65 * 320 * 4 + 65 * 4096 * 4 + 320 * 4096 * 4)
for rnd in range(2):
    for s in ("tf32", "f16"):
        set_split(s)
        bu = burst(); time.sleep(1.0)
        su = sustained()
        print(f"round {rnd} split {s:5s}: burst {bu*1e3:7.1f} us ({bytes_/bu/1e6:6.0f} GB/s)   sustained {su*1e3:7.1f} us ({bytes_/su/1e6:6.0f} GB/s)", flush=True)
        time.sleep(2.0)
# accuracy on 8 images against float64
sl = slice(0, 8)
for s in ("tf32", "f16"):
    set_split(s)
    call(); torch.cuda.synchronize()
    ref = torch.einsum("nkp,nkc->ncp", comp[sl].double().flatten(2), f[sl].double()).view(8, 320, 64, 64)
    err = (grid[sl].double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"split {s}: max err / scale vs float64 contraction of the kernel's own maps = {err:.2e}")
set_split("f16")
