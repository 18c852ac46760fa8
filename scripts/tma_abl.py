"""Ablations of the TMA stage-3 kernel (BLOBSPLAT_ST_ABL bits, measurement only): which engine bounds an item?"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from blobctrl_b200 import ops
    g = torch.Generator().manual_seed(1)
    def make(n, k, s, c):
        sc = torch.rand(n, k, s, s, generator=g)
        return (sc / sc.sum(1, keepdim=True)).cuda().to(torch.bfloat16), torch.randn(n, k, c, generator=g).cuda().to(torch.bfloat16)
    def timed(fn, reps=30, warm=5):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    res = []
    for name, n, k, lv in (("levels", 64, 33, [(32, 640), (16, 1280), (8, 1280)]), ("lvl64", 64, 33, [(64, 320)]), ("n1024", 1024, 65, [(64, 320)])):
        scs, fts = zip(*[make(n, k, s, c) for s, c in lv])
        outs = ops.feature_splat_levels(list(scs), list(fts), engine="tma")
        ptr = [o for o in outs]
        import ctypes
        from blobctrl_b200 import _capi as C
        res.append(f"{name} {timed(lambda: ops.feature_splat_levels(list(scs), list(fts), engine='tma')):7.1f} us")
    print(os.environ.get("BLOBSPLAT_ST_ABL", "0").rjust(3), " | ".join(res))
else:
    for mode in ("direct", "tma"):
        print("store mode:", mode, flush=True)
        for abl in (0, 1, 2, 4, 8, 9, 15):
            env = dict(os.environ, BLOBSPLAT_ST_ABL=str(abl), BLOBSPLAT_ST_STORE=mode)
            subprocess.run([sys.executable, __file__, "child"], env=env)
