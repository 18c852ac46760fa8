"""Where do the ~9 us of a one-tile launch go?  Needs a -DBS_TIMING=1 build (argv[1]); prints clock64 deltas of CTA 0."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blobctrl_b200 import _capi as C
L = ctypes.CDLL(sys.argv[1])
L.blobsplat_feature_splat.argtypes = C.SIGNATURES["blobsplat_feature_splat"]; L.blobsplat_feature_splat.restype = ctypes.c_int
names = ["entry", "tmem+barriers", "operands staged", "weights in stash", "A in TMEM", "D half0 ready", "D half1 ready",
         "half0 drained", "half1 drained", "all warps done"]
for (n, k, s, c, dt, code) in ((1, 33, 8, 320, torch.bfloat16, 2), (1, 33, 64, 320, torch.bfloat16, 2), (1, 33, 8, 320, torch.float32, 0)):
    sc = torch.rand(n, k, s, s, device="cuda"); sc = (sc / sc.sum(1, keepdim=True)).to(dt)
    ft = torch.randn(n, k, c, device="cuda").to(dt)
    out = torch.empty(n, c, s, s, device="cuda", dtype=dt)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: L.blobsplat_feature_splat(sc.data_ptr(), k * s * s, s * s, 1, ft.data_ptr(), out.data_ptr(), n, k, c, s, s, code, 2, 0, st)
    for _ in range(5): call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): call()
    b.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    L.blobsplat_debug_timing(buf)
    t0 = buf[0]
    print(f"N={n} K={k} {s}x{s} C={c} {dt}: {a.elapsed_time(b) / 50 * 1e3:.1f} us per back-to-back launch; CTA 0 (us at 1.9 GHz):")
    print("   " + ", ".join(f"{nm} {(buf[i] - t0) / 1900:.2f}" for i, nm in enumerate(names)))
