"""Where the bench's e2e step spends its time above the raw H2D: HostRenderer alone / + sample D2H on the main stream /
+ sample D2H on a side stream, each with and without the nvidia-smi clock sampler running (bench.ClockSampler)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic, ClockSampler, time_steps
from blobctrl_b200.streaming import HostRenderer
blobs, feats = synthetic(1024, 64, 320, seed=0)
pin = {k: v.pin_memory() for k, v in blobs.items()}; pf = feats.pin_memory()
r = HostRenderer(1024, 64, 64, 320, torch.float32, "cuda", chunks=4)
out_host = torch.empty((65 + 320, 64, 64), dtype=torch.float32).pin_memory()
side = torch.cuda.Stream(); ev = torch.cuda.Event()
def plain():
    r(pin["xs"], pin["ys"], pin["covs"], pin["sizes"], pf)
def d2h_main():
    o = r(pin["xs"], pin["ys"], pin["covs"], pin["sizes"], pf)
    out_host[:65].copy_(o["scores_pyramid"][64][-1], non_blocking=True)
    out_host[65:].copy_(o["feature_grid"][-1], non_blocking=True)
def d2h_side():
    o = r(pin["xs"], pin["ys"], pin["covs"], pin["sizes"], pf)
    ev.record(); side.wait_event(ev)
    with torch.cuda.stream(side):
        out_host[:65].copy_(o["scores_pyramid"][64][-1], non_blocking=True)
        out_host[65:].copy_(o["feature_grid"][-1], non_blocking=True)
    torch.cuda.current_stream().wait_stream(side) if False else None
for rnd in range(2):
    for name, fn in (("HostRenderer alone", plain), ("+ sample D2H, main stream", d2h_main), ("+ sample D2H, side stream", d2h_side)):
        t = time_steps(fn, 20, 5, lambda: None) / 20 * 1e3
        with ClockSampler(0) as c:
            ts = time_steps(fn, 20, 5, lambda: None) / 20 * 1e3
        torch.cuda.synchronize()
        print(f"{name:32s} {t:.3f} ms   with nvidia-smi -lms 20 sampling: {ts:.3f} ms   {c.summary()}", flush=True)
