"""How close is the host-input path to the PCIe ceiling?  Raw pinned H2D of the step's inputs vs HostRenderer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic
from blobctrl_b200.streaming import HostRenderer
blobs, feats = synthetic(1024, 64, 320, seed=0)
pin = {k: v.pin_memory() for k, v in blobs.items()}; pf = feats.pin_memory()
dev_f = torch.empty_like(feats, device="cuda")
def timed(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
t = timed(lambda: dev_f.copy_(pf, non_blocking=True))
print(f"raw H2D of the features ({pf.numel()*4/1e6:.1f} MB): {t:.3f} ms = {pf.numel()*4/t/1e6:.1f} GB/s")
for sets in (2,):
    for chunks in (1, 2, 3, 4, 6, 8):
        r = HostRenderer(1024, 64, 64, 320, torch.float32, "cuda", chunks=chunks, input_sets=sets)
        print(f"HostRenderer input_sets={sets} chunks={chunks}: {timed(lambda: r(pin['xs'], pin['ys'], pin['covs'], pin['sizes'], pf)):.3f} ms")
