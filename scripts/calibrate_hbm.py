"""Calibrate what this B200 sustains for the access mixes that matter here (CUDA events, best of 10):
copy (read+write, the MEASURED_PEAKS method), write-only fill, and read-only sum."""
import torch, json
x = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")   # 2 GiB
y = torch.empty_like(x)
def best(fn, nbytes, reps=10):
    ts = []
    for _ in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return nbytes / (min(ts[2:]) * 1e-3) / 1e9
res = {"copy_rw_GBs": best(lambda: y.copy_(x), 2 * x.numel() * 2),
       "fill_write_only_GBs": best(lambda: y.fill_(1.0), x.numel() * 2),
       "zero_write_only_GBs": best(lambda: y.zero_(), x.numel() * 2),
       "sum_read_only_GBs": best(lambda: x.view(torch.int16).sum(), x.numel() * 2)}
print(json.dumps(res))
