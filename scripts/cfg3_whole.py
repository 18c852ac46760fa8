"""cfg3 end to end (64 images, 32 blobs, 64/32/16/8, 320..1280 channels, bf16): eager and graph time, roofline fraction."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blobctrl_b200 as B
from bench import synthetic
dev = "cuda"
g = torch.Generator().manual_seed(1)
hb, _ = synthetic(64, 32, 1, seed=0)
blobs = {kk: v.to(dev) for kk, v in hb.items()}
chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
lf = {s: torch.randn(64, 33, c, generator=g).to(dev).to(torch.bfloat16) for s, c in chans.items()}
by3 = 64 * (28 * 32 + sum(33 * c * 2 + 33 * s * s * 2 + c * s * s * 2 for s, c in chans.items()))
def timed(fn, reps=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
fn = lambda: B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16)
for rnd in range(3):
    t_e = timed(fn)
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = fn()
    t_g = timed(gr.replay)
    print(f"cfg3 whole: eager {t_e:6.1f} us  graph {t_g:6.1f} us   frac of 6542.7 GB/s: eager {by3 / t_e / 1e3 / 6542.7:.3f} graph {by3 / t_g / 1e3 / 6542.7:.3f}", flush=True)
import time
t0 = time.perf_counter()
for _ in range(200): fn()
host = (time.perf_counter() - t0) / 200 * 1e6
torch.cuda.synchronize()
print(f"host issue time per call (async): {host:.1f} us")
