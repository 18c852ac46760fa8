"""A/B builds of the library (BLOBSPLAT_LIB) on BASELINE config 3 (one call: render + pyramid, then the TMA engine) and on its lower levels alone."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import blobctrl_b200 as B
    from blobctrl_b200 import ops
    from bench import synthetic
    def t(fn, reps=100, warm=10):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    hb, _ = synthetic(64, 32, 8, seed=0)
    blobs = {k: v.cuda() for k, v in hb.items()}
    g = torch.Generator().manual_seed(1)
    lf = {s: torch.randn(64, 33, c, generator=g).cuda().to(torch.bfloat16) for s, c in ((64, 320), (32, 640), (16, 1280), (8, 1280))}
    call = lambda: B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16)
    out = call(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): keep = call()
    scs = [out["scores_pyramid"][s] for s in (32, 16, 8)]; fts = [lf[s] for s in (32, 16, 8)]
    lv = lambda: ops.feature_splat_levels(scs, fts, engine="tma")
    lv(); torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2): keep2 = lv()
    print(f"{os.path.basename(os.environ.get('BLOBSPLAT_LIB', 'default')):22s} cfg3 eager {t(call):6.2f} us, graph {t(gr.replay):6.2f} us | lower levels graph {t(g2.replay):6.2f} us", flush=True)
else:
    for rnd in range(3):
        for lib in sys.argv[1:]:
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, BLOBSPLAT_LIB=os.path.abspath(lib)))
