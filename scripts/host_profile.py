import sys, os, cProfile, pstats, time
sys.path.insert(0, "/root/repo")
import torch
from bench import synthetic
import blobctrl_b200 as B
hb, _ = synthetic(64, 32, 1, seed=0)
blobs = {k: v.cuda() for k, v in hb.items()}
chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
g = torch.Generator().manual_seed(1)
lf = {s: torch.randn(64, 33, c, generator=g).cuda().to(torch.bfloat16) for s, c in chans.items()}
fn = lambda: B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16)
for _ in range(20): fn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): fn()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host issue time per call (us):", (t1 - t0) / 200 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): fn()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
