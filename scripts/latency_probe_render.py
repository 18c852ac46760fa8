"""Phases of a cfg2-sized fused render (1 image, 16 blobs, 64x64, C = 320): clock64 stamps of CTA 0, -DBS_TIMING=1 build (argv[1])."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blobctrl_b200 import _capi as C, ops
from bench import synthetic
L = ctypes.CDLL(sys.argv[1])
L.blobsplat_render.argtypes = C.SIGNATURES["blobsplat_render"]; L.blobsplat_render.restype = ctypes.c_int
names = ["entry", "tmem+barriers", "operands staged", "weights in stash", "A in TMEM", "D part0 ready", "D part1 ready",
         "part0 drained", "part1 drained", "all warps done"]
for (n, m, c, dt, code) in ((1, 16, 320, torch.float32, 0), (1, 16, 64, torch.float32, 0), (1, 64, 320, torch.float32, 0)):
    hb, hf = synthetic(n, m, c, seed=0)
    b = {k: v.cuda() for k, v in hb.items()}
    xs, ys, covs, sizes, _, _ = ops.canonical_blobs(**b)
    f = hf.cuda().to(dt)
    comp = torch.empty(n, m + 1, 64, 64, device="cuda", dtype=dt); grid = torch.empty(n, c, 64, 64, device="cuda", dtype=dt)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: L.blobsplat_render(xs.data_ptr(), ys.data_ptr(), covs.data_ptr(), sizes.data_ptr(), f.data_ptr(), code, n, m, 64, 64, c,
                                      comp.data_ptr(), grid.data_ptr(), code, 0, st)
    for _ in range(5): assert call() == 0
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): call()
    e.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    L.blobsplat_debug_timing_render(buf)
    t0 = buf[0]
    print(f"N={n} M={m} C={c} {dt}: {a.elapsed_time(e) / 50 * 1e3:.1f} us per back-to-back launch; CTA 0 (clock64 / 1900), last of the back-to-back "
          f"launches (its head overlaps the previous launch's tail: programmatic dependent launch):")
    print("   " + ", ".join(f"{nm} {(buf[i] - t0) / 1900:.2f}" for i, nm in enumerate(names)))
    torch.cuda.synchronize(); call(); torch.cuda.synchronize()
    L.blobsplat_debug_timing_render(buf)
    t0 = buf[0]
    print("   alone (idle GPU before the launch): " + ", ".join(f"{nm} {(buf[i] - t0) / 1900:.2f}" for i, nm in enumerate(names)))
