"""BASELINE config 2 (1 image, 16 blobs, 64x64, C = 320, float32) a few times — the `ncu` target for the latency kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blobctrl_b200 as B
from bench import synthetic
hb, hf = synthetic(1, 16, 320, seed=0)
b = {k: v.cuda() for k, v in hb.items()}; f = hf.cuda()
for _ in range(4):
    out = B.splat_features(**b, features=f, score_size=64, interp_size=64, ret_layout=False)
torch.cuda.synchronize()
print("done", out["feature_grid"].shape)
