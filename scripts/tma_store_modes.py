"""TMA stage-3 engine: drain store path A/B (BLOBSPLAT_ST_STORE = direct | tma | hybrid), one process per mode, interleaved."""
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
    def timed(fn, reps=50, warm=5):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    res = []
    for name, n, k, lv in (("levels", 64, 33, [(32, 640), (16, 1280), (8, 1280)]), ("4levels", 64, 33, [(64, 320), (32, 640), (16, 1280), (8, 1280)]),
                           ("n1024", 1024, 65, [(64, 320)])):
        scs, fts = zip(*[make(n, k, s, c) for s, c in lv])
        fn = lambda: ops.feature_splat_levels(list(scs), list(fts), engine="tma")
        outs = fn(); torch.cuda.synchronize()
        ref = torch.einsum("nkhw,nkc->nchw", scs[0][:2].double(), fts[0][:2].double())
        err = ((outs[0][:2].double() - ref).abs().max() / ref.abs().max()).item()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            keep = fn()
        res.append(f"{name} {timed(gr.replay):7.1f} us (err {err:.1e})")
    print(os.environ.get("BLOBSPLAT_ST_STORE", "-").rjust(7), " | ".join(res), flush=True)
else:
    for rnd in range(2):
        for mode in ("direct", "tma", "hybrid"):
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, BLOBSPLAT_ST_STORE=mode))
