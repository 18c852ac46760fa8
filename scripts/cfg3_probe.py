"""cfg3 (multi-scale bf16): eager wall vs CUDA-graph replay vs per-kernel device time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic
import blobctrl_b200 as B
hb, _ = synthetic(64, 32, 1, seed=0)
blobs = {k: v.cuda() for k, v in hb.items()}
chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
g = torch.Generator().manual_seed(1)
lf = {s: torch.randn(64, 33, c, generator=g).cuda().to(torch.bfloat16) for s, c in chans.items()}
fn = lambda: B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16)
def timed(f, reps=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
print("eager us:", timed(fn))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fn(); 
torch.cuda.current_stream().wait_stream(s)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    out = fn()
print("graph replay us:", timed(gr.replay))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): fn()
    torch.cuda.synchronize()
for e in prof.key_averages():
    if e.device_time_total > 0: print(f"{e.key[:70]:70s} n={e.count} dev_us_avg={e.device_time_total/e.count:.1f}")
