"""Launch the fused render (cfg5b shape, fewer images) a few times — the target of `ncu` captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic
from blobctrl_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
blobs, feats = synthetic(n, 64, 320, seed=0)
b = {k: v.cuda() for k, v in blobs.items()}
f = feats.cuda().to(dtype)
for _ in range(3):
    d, g = ops.render_fused(**b, features=f, height=64, width=64, out_dtype=dtype)
torch.cuda.synchronize()
print("done", d.shape, g.shape)
