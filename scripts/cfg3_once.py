import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blobctrl_b200 as B
from bench import synthetic
g = torch.Generator().manual_seed(1)
hb, _ = synthetic(64, 32, 1, seed=0)
blobs = {kk: v.cuda() for kk, v in hb.items()}
chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
lf = {s: torch.randn(64, 33, c, generator=g).cuda().to(torch.bfloat16) for s, c in chans.items()}
for _ in range(4):
    B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
