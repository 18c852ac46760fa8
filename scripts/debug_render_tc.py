"""GPU debug harness for the fused tcgen05 render: compares against the FMA path / fp64 oracle and prints
where and how the output differs (used while bringing the kernel up; not part of the test-suite)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from blobctrl_b200 import ops
from oracle import blob_oracle

def run(n, m, s, c, dtype=torch.float32, seed=0, ident=False):
    syn = blob_oracle.synthetic_blobs(n, m, seed=seed, c=c)
    b = {k: torch.from_numpy(v).cuda() for k, v in syn.items() if k != "features"}
    f = torch.from_numpy(syn["features"]).cuda()
    if ident:   # features[k, c] = 1 if c == k: grid channel c == composed plane c
        f.zero_()
        for k in range(min(m + 1, c)):
            f[:, k, k] = 1
    f = f.to(dtype)
    d_ref, _ = ops.render_scores(**b, height=s, width=s, out_dtype=dtype)
    g_ref = ops.feature_splat(d_ref, f)
    torch.cuda.synchronize()
    d, g = ops.render_fused(**b, features=f, height=s, width=s, out_dtype=dtype)
    torch.cuda.synchronize()
    ed = (d.float() - d_ref.float()).abs().max().item()
    eg = (g.float() - g_ref.float()).abs()
    scale = g_ref.float().abs().max().item()
    print(f"N={n} M={m} S={s} C={c} {dtype} ident={ident}: composed err {ed:.3e}; grid err {eg.max().item():.3e} (scale {scale:.3g})")
    if eg.max().item() > (1e-5 if dtype == torch.float32 else 2e-2) * scale:
        bad = (eg > 1e-4 * scale)
        print("  bad fraction", bad.float().mean().item())
        print("  bad per image", bad.float().mean((1, 2, 3)).cpu().numpy()[:8])
        bc = bad.float().mean((0, 2, 3)).cpu().numpy()
        print("  bad channels (first 48)", np.round(bc[:48], 2))
        print("  bad channel idx", np.nonzero(bc > 0)[0][:64])
        bp = bad.float().mean((0, 1)).reshape(-1).cpu().numpy()
        print("  bad pixels idx (first 40)", np.nonzero(bp > 0)[0][:40], "count", (bp > 0).sum())
        print("  sample got/ref", g[0, :4, 0, :4].float().cpu().numpy(), g_ref[0, :4, 0, :4].float().cpu().numpy())
        return False
    return True

if __name__ == "__main__":
    ok = True
    ok &= run(1, 7, 16, 32, ident=True)
    ok &= run(1, 7, 16, 32)
    ok &= run(2, 16, 64, 320)
    ok &= run(3, 64, 64, 320)
    ok &= run(2, 32, 64, 640)
    ok &= run(2, 64, 64, 320, torch.bfloat16)
    ok &= run(2, 33, 32, 1280, torch.float16)
    ok &= run(5, 20, 24, 96)
    print("ALL OK" if ok else "FAILURES")
