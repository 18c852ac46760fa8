"""Where the host time of one small call goes (cfg2: N=1, 16 blobs, 64x64, C=320, fp32): each piece in a loop of its own."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic
import blobctrl_b200 as B
from blobctrl_b200 import ops, _capi as C
hb, hf = synthetic(1, 16, 320, seed=0)
blobs = {k: v.cuda() for k, v in hb.items()}
f = hf.cuda()
def per_call(fn, n=3000):
    for _ in range(100): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6
xs, ys, covs, sizes, n, m = ops.canonical_blobs(**blobs)
comp = torch.empty((1, 17, 64, 64), device="cuda"); grid = torch.empty((1, 320, 64, 64), device="cuda")
fn_c = C.lib().blobsplat_render
args = (xs.data_ptr(), ys.data_ptr(), covs.data_ptr(), sizes.data_ptr(), f.data_ptr(), 0, 1, 16, 64, 64, 320, comp.data_ptr(), grid.data_ptr(), 0, 0, None)
print(f"C call alone (blobsplat_render, fixed arguments)      {per_call(lambda: fn_c(*args)):6.2f} us")
print(f"torch.empty x2                                         {per_call(lambda: (torch.empty((1, 17, 64, 64), device='cuda'), torch.empty((1, 320, 64, 64), device='cuda'))):6.2f} us")
print(f"canonical_blobs                                        {per_call(lambda: ops.canonical_blobs(**blobs)):6.2f} us")
print(f"stream_of + dev_of                                     {per_call(lambda: (C.stream_of(covs), C.dev_of(covs))):6.2f} us")
print(f"ops.render_fused                                       {per_call(lambda: ops.render_fused(**blobs, features=f, height=64, width=64)):6.2f} us")
print(f"splat_features (ret_layout=False)                      {per_call(lambda: B.splat_features(**blobs, features=f, score_size=64, interp_size=64, ret_layout=False)):6.2f} us")
print(f"splat_features(return_d_score=True), 512x512, 1 blob   {per_call(lambda: B.splat_features(xs[:, :1], ys[:, :1], covs[:, :1], sizes[:, :1], score_size=(512, 512), return_d_score=True)):6.2f} us")
