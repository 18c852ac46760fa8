"""TMA stage-3 engine (splat_tma.cu): correctness against a float64 contraction of the same 16-bit inputs on a sweep of
shapes (ragged channel / pixel tails, one and several levels, strided score views), then timings against the other engines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blobctrl_b200 import ops

dev = "cuda"
g = torch.Generator().manual_seed(3)


def make(n, k, hw, c, dt):
    h, w = hw
    sc = torch.rand(n, k, h, w, generator=g)
    sc = (sc / sc.sum(1, keepdim=True)).to(dev).to(dt)
    ft = torch.randn(n, k, c, generator=g).to(dev).to(dt)
    return sc, ft


def ref(sc, ft):
    return torch.einsum("nkhw,nkc->nchw", sc.double(), ft.double())


bad = 0
cases = [  # (n, k, [(hw, c), ...])
    (2, 33, [((32, 32), 640), ((16, 16), 1280), ((8, 8), 1280)]),
    (3, 17, [((64, 64), 320)]),
    (1, 1, [((64, 64), 1024)]),
    (2, 65, [((64, 64), 320), ((32, 32), 640)]),
    (2, 5, [((24, 24), 72), ((12, 12), 136), ((4, 4), 8)]),
    (1, 128, [((48, 40), 200)]),
    (5, 33, [((8, 8), 64), ((4, 2), 1280)]),
    (2, 40, [((20, 18), 328), ((10, 12), 96), ((64, 64), 320), ((2, 4), 64)]),
]
for dt, tol in ((torch.bfloat16, 1e-2), (torch.float16, 2e-3)):
    for n, k, lv in cases:
        scs, fts = zip(*[make(n, k, hw, c, dt) for hw, c in lv])
        outs = ops.feature_splat_levels(list(scs), list(fts), engine="tma")
        torch.cuda.synchronize()
        for (hw, c), sc, ft, o in zip(lv, scs, fts, outs):
            want = ref(sc, ft)
            err = ((o.double() - want).abs().max() / want.abs().max()).item()
            # rounding only: the fp32-accumulated sum rounded once to 16 bits
            exact = torch.equal(o, want.to(dt))
            other = ops.feature_splat(sc, ft, engine="tensor") if k - 1 <= 127 else None
            same = None if other is None else torch.equal(o, other)
            ok = err <= tol
            bad += 0 if ok else 1
            print(f"{str(dt)[6:]:9s} n={n} k={k} {hw} c={c}: err/scale {err:.2e} {'ok' if ok else 'BAD'}  == rounded f64: {exact}  == tensor engine: {same}")
    # strided score views (a slice of a larger buffer: plane stride > P)
    big = torch.rand(2, 20, 40, 32, generator=g).to(dev).to(dt)
    view = big[:, 2:19, :32, :]            # not pixel-linear? rows contiguous, 32x32 of 40x32 -> linear (first 1024 px of each plane)
    ft = torch.randn(2, 17, 256, generator=g).to(dev).to(dt)
    o = ops.feature_splat(view, ft, engine="tma")
    want = ref(view, ft)
    err = ((o.double() - want).abs().max() / want.abs().max()).item()
    bad += 0 if err <= tol else 1
    print(f"{str(dt)[6:]:9s} strided view: err/scale {err:.2e}")
print("FAILURES:", bad)


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


def bench(name, n, k, lv, dt=torch.bfloat16, engines=("auto", "tensor", "tma")):
    scs, fts = zip(*[make(n, k, hw, c, dt) for hw, c in lv])
    scs, fts = list(scs), list(fts)
    byt = n * sum(k * c * 2 + k * h * w * 2 + c * h * w * 2 for (h, w), c in lv)
    for eng in engines:
        try:
            if len(lv) == 1:
                fn = lambda: ops.feature_splat(scs[0], fts[0], engine=eng)
            else:
                fn = lambda: ops.feature_splat_levels(scs, fts, engine=eng)
            t_e = timed(fn)
            gr, keep = graphed(fn)
            t_g = timed(gr.replay)
            print(f"{name:28s} {eng:7s}: eager {t_e:7.1f} us  graph {t_g:7.1f} us  {byt / t_g / 1e3:6.0f} GB/s = {byt / t_g / 1e3 / 6542.7:.3f}")
        except Exception as e:
            print(f"{name:28s} {eng:7s}: {type(e).__name__}: {str(e)[:100]}")


bench("cfg3 levels 32/16/8", 64, 33, [((32, 32), 640), ((16, 16), 1280), ((8, 8), 1280)])
bench("cfg3 levels 64/32/16/8", 64, 33, [((64, 64), 320), ((32, 32), 640), ((16, 16), 1280), ((8, 8), 1280)])
bench("cfg3 level 64 alone", 64, 33, [((64, 64), 320)])
bench("cfg5c stage 3 (1024 img)", 1024, 65, [((64, 64), 320)])
bench("pipeline K=1 C=1024 x16", 16, 1, [((64, 64), 1024)])
bench("cfg2-like N=1 K=17", 1, 17, [((64, 64), 320)])
