"""A/B builds of the one-image latency kernel (BLOBSPLAT_LIB): graph replays of single small renders, one process per build."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from bench import synthetic
    from blobctrl_b200 import ops
    def t(fn, reps=400, warm=20):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    res = []
    for (m, s, c) in ((16, 64, 320), (1, 64, 1024), (8, 128, 64), (32, 64, 160), (1, 512, 3)):
        hb, hf = synthetic(1, m, c, seed=0)
        bb = {k: v.cuda() for k, v in hb.items()}; ff = hf.cuda()
        ops.render_small(**bb, features=ff, height=s, width=s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): keep = ops.render_small(**bb, features=ff, height=s, width=s)
        res.append(f"M={m} S={s} C={c}: {t(g.replay):5.2f}")
    e = torch.cuda.CUDAGraph()
    x = torch.zeros(1, device="cuda")
    with torch.cuda.graph(e): x.add_(1)
    res.append(f"one tiny elementwise kernel: {t(e.replay):5.2f}")
    print(f"{os.path.basename(os.environ.get('BLOBSPLAT_LIB', 'default')):16s} " + " | ".join(res) + "  (us per graph replay)", flush=True)
else:
    for rnd in range(2):
        for lib in sys.argv[1:]:
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, BLOBSPLAT_LIB=os.path.abspath(lib)))
