"""Timeline of CTA 0 of the TMA stage-3 kernel (clock64 stamps, BLOBSPLAT_ST_DBG_PTR): where does an item's time go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dbg = torch.zeros(4 * 64 * 4 + 2 * 160, dtype=torch.int64, device="cuda")
if len(sys.argv) > 2:
    os.environ["BLOBSPLAT_ST_DBG_CTA"] = sys.argv[2]
os.environ["BLOBSPLAT_ST_DBG_PTR"] = str(dbg.data_ptr())
from blobctrl_b200 import ops
g = torch.Generator().manual_seed(1)
def make(n, k, s, c):
    sc = torch.rand(n, k, s, s, generator=g)
    return (sc / sc.sum(1, keepdim=True)).cuda().to(torch.bfloat16), torch.randn(n, k, c, generator=g).cuda().to(torch.bfloat16)
which = sys.argv[1] if len(sys.argv) > 1 else "n1024"
n, k, lv = {"levels": (64, 33, [(32, 640), (16, 1280), (8, 1280)]), "lvl64": (64, 33, [(64, 320)]), "n1024": (1024, 65, [(64, 320)])}[which]
scs, fts = zip(*[make(n, k, s, c) for s, c in lv])
for _ in range(3):
    ops.feature_splat_levels(list(scs), list(fts), engine="tma")
torch.cuda.synchronize()
span = dbg.cpu()[1024:1024 + 296].view(148, 2)
d = dbg.cpu()[:1024].view(4, 64, 4)
t0 = int(d[1, 0, 0])
names = ["producer: start / f_free ok / issued", "MMA: start / operands ok / slot ok / committed", "drain h0: wait / d_full / released / done", "drain h1: wait / d_full / released / done"]
t_first = int(span[:, 0].min())
print("per-CTA (start, end) in us since the first CTA started:")
print(" ".join(f"{(int(a) - t_first) / 1e3:.1f}-{(int(b) - t_first) / 1e3:.1f}" for a, b in span[::6]))
print("latest end:", (int(span[:, 1].max()) - t_first) / 1e3, "us; CTA", int(span[:, 1].argmax()))
for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 24):
    row = []
    for r in range(4):
        row.append(" ".join(f"{(int(x) - t0):7d}" if int(x) else "      -" for x in d[r, i]))
    print(f"{i:3d} | P {row[0][:23]} | M {row[1]} | D0 {row[2]} | D1 {row[3]}")
print("(ns since the MMA thread first item, globaltimer)")
