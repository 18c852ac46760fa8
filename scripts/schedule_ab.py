"""A/B of the fused render's two work schedules (whole runs round-robin vs equal tile ranges) at the strong-scaling shard
sizes of BASELINE configs[4]: 1024/N images per GPU.  BLOBSPLAT_TC_SCHEDULE is read per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic
from blobctrl_b200 import ops


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for dtype in (torch.float32, torch.bfloat16):
    for n in (1024, 512, 256, 128, 64):
        blobs, feats = synthetic(n, 64, 320, seed=0)
        b = {k: v.cuda() for k, v in blobs.items()}
        f = feats.cuda().to(dtype)
        comp = torch.empty((n, 65, 64, 64), dtype=dtype, device="cuda"); grid = torch.empty((n, 320, 64, 64), dtype=dtype, device="cuda")
        res = {}
        for sched in ("auto", "whole", "ranges"):
            if sched == "auto":
                os.environ.pop("BLOBSPLAT_TC_SCHEDULE", None)
            else:
                os.environ["BLOBSPLAT_TC_SCHEDULE"] = sched
            res[sched] = timed(lambda: ops.render_fused_into(b["xs"], b["ys"], b["covs"], b["sizes"], f, 64, 64, comp, grid))
        os.environ.pop("BLOBSPLAT_TC_SCHEDULE", None)
        ideal = res["auto"] if n == 1024 else None
        print(f"{str(dtype):16s} N={n:5d}  auto {res['auto']*1e3:8.1f} us  whole {res['whole']*1e3:8.1f} us  ranges {res['ranges']*1e3:8.1f} us"
              f"   us/image: {res['auto']*1e3/n:.3f}")
