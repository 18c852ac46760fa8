"""Bandwidth of the small 'next row' kernels (residual injection, conditioning fill, pyramid) at SD-1.5 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blobctrl_b200 as B
from blobctrl_b200.pipelines.conditioning import inject_residual, BlobNetInputBuffers
from blobctrl_b200 import ops
def timed(f, reps=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
dt = torch.float16
for (b, c, h) in ((16, 320, 64), (16, 640, 32), (16, 1280, 16), (16, 1280, 8), (64, 320, 64)):
    hid = torch.randn(b, c, h, 2 * h, device="cuda", dtype=dt)      # UNet hidden state on the doubled-width canvas
    res = torch.randn(b, c, h, 2 * h, device="cuda", dtype=dt)
    us = timed(lambda: inject_residual(hid, res, 1.0))
    byts = b * c * h * h * 2 * 3                                      # right half: read hidden + residual, write hidden
    print(f"residual_inject B={b} C={c} {h}x{2*h} f16: {us:.1f} us = {byts / us / 1e3:.0f} GB/s")
for (n, k, s) in ((64, 33, 64), (1024, 65, 64)):
    d = torch.rand(n, k, s, s, device="cuda").to(torch.bfloat16)
    us = timed(lambda: ops.halving_pyramid(d, 3))
    byts = n * k * s * s * 2 * (1 + 1 / 4 + 1 / 16 + 1 / 64)
    print(f"pyramid N={n} K={k} {s}->8 bf16: {us:.1f} us = {byts / us / 1e3:.0f} GB/s")
for bsz in (2, 16):
    buf = BlobNetInputBuffers(bsz, 64, 64, 1024, dt)
    fg = torch.rand(bsz, 1, 64, 64, device="cuda", dtype=dt); bg = torch.rand_like(fg)
    ff = torch.randn(bsz, 1, 1024, device="cuda", dtype=dt); lat = torch.randn(bsz, 4, 64, 64, device="cuda", dtype=dt)
    try:
        us = timed(lambda: buf.fill_static(fg, bg, ff, lat, lat))
        byts = bsz * 1029 * 64 * 128 * 2
        print(f"conditioning fill_static B={bsz}: {us:.1f} us = {byts / us / 1e3:.0f} GB/s of canvas writes")
    except Exception as e:
        print("fill_static probe failed:", str(e)[:120])
