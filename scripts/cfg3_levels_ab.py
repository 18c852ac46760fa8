"""cfg3's three lower pyramid levels (32/16/8; 640/1280/1280 channels; 64 images; bf16): per-level launches vs ONE launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blobctrl_b200 import ops
import blobctrl_b200 as B
from bench import synthetic

dev = "cuda"
n, k = 64, 33
levels = [(32, 640), (16, 1280), (8, 1280)]
g = torch.Generator().manual_seed(1)
scs, fts = [], []
for s, c in levels:
    sc = torch.rand(n, k, s, s, generator=g)
    scs.append((sc / sc.sum(1, keepdim=True)).to(dev).to(torch.bfloat16))
    fts.append(torch.randn(n, k, c, generator=g).to(dev).to(torch.bfloat16))


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def graphed(fn):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = fn()
    return gr, keep


per = ops.feature_splat_levels(scs, fts)
one = ops.feature_splat_levels(scs, fts, engine="tensor")
for (s, c), a, b in zip(levels, one, per):
    print(f"level {s} C={c}: one-launch vs per-level max diff {(a.float() - b.float()).abs().max().item():.3e}  equal {torch.equal(a, b)}")
byt = n * sum(k * c * 2 + k * s * s * 2 + c * s * s * 2 for s, c in levels)
for name, eng in (("per-level", "auto"), ("one launch", "tensor")):
    t_e = timed(lambda: ops.feature_splat_levels(scs, fts, engine=eng))
    gr, keep = graphed(lambda: ops.feature_splat_levels(scs, fts, engine=eng))
    t_g = timed(gr.replay)
    print(f"{name:11s}: eager {t_e:6.1f} us   graph {t_g:6.1f} us   ({byt / t_g / 1e3:6.0f} GB/s in the graph)")
# whole cfg3
hb, _ = synthetic(64, 32, 1, seed=0)
blobs = {kk: v.to(dev) for kk, v in hb.items()}
chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
lf = {s: torch.randn(64, 33, c, generator=g).to(dev).to(torch.bfloat16) for s, c in chans.items()}
by3 = 64 * (28 * 32 + sum(33 * c * 2 + 33 * s * s * 2 + c * s * s * 2 for s, c in chans.items()))
fn = lambda: B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16)
t_e = timed(fn)
gr, keep = graphed(fn)
t_g = timed(gr.replay)
print(f"cfg3 whole: eager {t_e:6.1f} us  graph {t_g:6.1f} us   frac of 6542.7 GB/s: eager {by3 / t_e / 1e3 / 6542.7:.3f} graph {by3 / t_g / 1e3 / 6542.7:.3f}")
