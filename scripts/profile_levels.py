"""cfg3's lower pyramid levels as one launch (bf16) a few times — the target of `ncu` captures of the multi-level kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blobctrl_b200 import ops

n, k = 64, 33
levels = [(32, 640), (16, 1280), (8, 1280)]
g = torch.Generator().manual_seed(1)
scs, fts = [], []
for s, c in levels:
    sc = torch.rand(n, k, s, s, generator=g)
    scs.append((sc / sc.sum(1, keepdim=True)).cuda().to(torch.bfloat16))
    fts.append(torch.randn(n, k, c, generator=g).cuda().to(torch.bfloat16))
eng = sys.argv[1] if len(sys.argv) > 1 else "tensor"
for _ in range(3):
    out = ops.feature_splat_levels(scs, fts, engine=eng)
torch.cuda.synchronize()
print("done", [o.shape for o in out])
