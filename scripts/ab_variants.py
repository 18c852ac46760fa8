"""Build compile-time variants of libblobsplat.so and A/B them in ONE process (interleaved rounds), so the
comparison is immune to box-to-box and clock drift.  Usage:
    python scripts/ab_variants.py build   (here, CPU)   ->  build/variants/<name>/libblobsplat.so
    python scripts/ab_variants.py run     (GPU box)
"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "r3_unit_scale": "-DBS_SKIP_UNIT_SCALE=0", "r3_skip_scale": "-DBS_SKIP_UNIT_SCALE=1",
}
VDIR = os.path.join(ROOT, "blobctrl_b200", "lib", "variants")

def build():
    for name, flags in VARIANTS.items():
        out = os.path.join(VDIR, name)
        if os.path.exists(os.path.join(out, "libblobsplat.so")) and "--force" not in sys.argv:
            continue
        os.makedirs(out, exist_ok=True)
        subprocess.run(["make", "-C", os.path.join(ROOT, "blobctrl_b200", "csrc"), "-j8", f"EXTRA={flags}",
                        f"OBJDIR={os.path.join(ROOT, 'build', 'variants', name)}", f"OUT={os.path.join(out, 'libblobsplat.so')}"],
                       check=True, stdout=subprocess.DEVNULL)
        print("built", name)

def run():
    import torch
    from bench import synthetic
    from blobctrl_b200 import _capi as C
    blobs, feats = synthetic(1024, 64, 320, seed=0)
    b = {k: v.cuda().contiguous() for k, v in blobs.items()}
    libs = {}
    names = [d for d in sorted(os.listdir(VDIR)) if os.path.exists(os.path.join(VDIR, d, "libblobsplat.so"))]
    if len(sys.argv) > 2:
        names = [n for n in names if n in sys.argv[2:]]
    for name in names:
        L = ctypes.CDLL(os.path.join(VDIR, name, "libblobsplat.so"))
        L.blobsplat_render.argtypes = C.SIGNATURES["blobsplat_render"]; L.blobsplat_render.restype = ctypes.c_int
        libs[name] = L
    for dtype, code in ((torch.float32, C.F32), (torch.bfloat16, C.BF16)):
        f = feats.cuda().to(dtype).contiguous()
        comp = torch.empty((1024, 65, 64, 64), dtype=dtype, device="cuda")
        grid = torch.empty((1024, 320, 64, 64), dtype=dtype, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        def call(L):
            rc = L.blobsplat_render(b["xs"].data_ptr(), b["ys"].data_ptr(), b["covs"].data_ptr(), b["sizes"].data_ptr(),
                                    f.data_ptr(), code, 1024, 64, 64, 64, 320, comp.data_ptr(), grid.data_ptr(), code, 0, st)
            assert rc == 0
        res = {n: [] for n in libs}
        for rnd in range(6):
            order = list(libs.items())
            order = order[rnd % len(order):] + order[:rnd % len(order)]     # rotate: no variant always runs first
            for n, L in order:
                for _ in range(3): call(L)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20): call(L)
                e1.record(); torch.cuda.synchronize()
                res[n].append(round(e0.elapsed_time(e1) / 20, 4))
        print(dtype)
        for n, v in res.items():
            call(libs[n]); torch.cuda.synchronize()      # stage-3 error of this variant against float64 on its own weights
            ref = torch.einsum("nkp,nkc->ncp", comp[:8].double().flatten(2), f[:8].double())
            err = ((grid[:8].double().flatten(2) - ref).abs().max() / ref.abs().max()).item()
            print(f"  {n:18s} {v}  median {sorted(v)[len(v)//2]}  stage-3 err/scale {err:.2e}")
        for tn in [n for n in libs if n.startswith("timing")]:
            call(libs[tn]); torch.cuda.synchronize()
            buf = (ctypes.c_ulonglong * 48)()
            libs[tn].blobsplat_debug_timing(buf)
            print(" ", tn)
            names = {0: "compute front (warp0)", 1: "compute back (warp4)", 2: "epilogue (first warp)", 3: "MMA warp"}
            for role in range(4):
                row = [buf[role * 12 + i] for i in range(12)]
                tot = sum(row) or 1
                print(f"  timing {names[role]:24s} total {tot/1e3:.0f}k cyc:", [f"{i}:{100*v/tot:.0f}%" for i, v in enumerate(row) if v])

if __name__ == "__main__":
    build() if sys.argv[1:2] == ["build"] else run()
