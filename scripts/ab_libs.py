"""A/B two builds of the library (BLOBSPLAT_LIB) on the headline shape and on the one-image launches, one process per build, interleaved.
    python scripts/ab_libs.py LIB_A LIB_B"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from bench import synthetic
    from blobctrl_b200 import ops
    import blobctrl_b200 as B
    def t(fn, reps=20, warm=3):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    blobs, feats = synthetic(1024, 64, 320, seed=0)
    b = {k: v.cuda() for k, v in blobs.items()}; f = feats.cuda()
    big = t(lambda: ops.render_fused(**b, features=f, height=64, width=64))
    sus = t(lambda: ops.render_fused(**b, features=f, height=64, width=64), reps=1200, warm=100)
    hb, hf = synthetic(1, 16, 320, seed=0)
    b2 = {k: v.cuda() for k, v in hb.items()}; f2 = hf.cuda()
    kw = dict(features=f2, score_size=64, interp_size=64, ret_layout=False)
    g = t(lambda: B.splat_features(**b2, **kw, cuda_graph=True), reps=300, warm=10) * 1e3
    hb3, hf3 = synthetic(4, 32, 320, seed=0)
    b3 = {k: v.cuda() for k, v in hb3.items()}; f3 = hf3.cuda()
    g3 = t(lambda: B.splat_features(**b3, features=f3, score_size=64, interp_size=64, ret_layout=False, cuda_graph=True), reps=300, warm=10) * 1e3
    print(f"{os.path.basename(os.environ.get('BLOBSPLAT_LIB', 'default')):28s} cfg5b burst {big:.4f} ms, 1200 back to back {sus:.4f} ms | cfg2 graph {g:.2f} us | 4 images x 32 blobs graph {g3:.2f} us", flush=True)
else:
    for rnd in range(2):
        for lib in sys.argv[1:]:
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, BLOBSPLAT_LIB=os.path.abspath(lib)))
