"""Host-side cost of one small call (cfg2: N=1, 16 blobs, 64x64, C=320, fp32)."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic
import blobctrl_b200 as B
hb, hf = synthetic(1, 16, 320, seed=0)
blobs = {k: v.cuda() for k, v in hb.items()}
f = hf.cuda()
fn = lambda: B.splat_features(**blobs, features=f, score_size=64, interp_size=64, ret_layout=False)
for _ in range(50): fn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(500): fn()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host issue time per call (us):", (t1 - t0) / 500 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(500): fn()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
