"""Host link at N ranks: ordinary pinned memory against write-combined pinned memory (cudaHostAllocWriteCombined).
torchrun --nproc-per-node N scripts/h2d_wc_probe.py  (or plain python for one rank).  All ranks copy at once; max over ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from blobctrl_b200.hostmem import pinned_empty
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
barrier = dist.barrier if world > 1 else (lambda: None)
shape = (1024, 65, 320)
src = torch.randn(shape)
dev = torch.empty(shape, device="cuda")
def timed(host, reps=20):
    for _ in range(3): dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize(); barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): dev.copy_(host, non_blocking=True)
    b.record(); torch.cuda.synchronize(); barrier()
    t = torch.tensor([a.elapsed_time(b) / reps], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
for wc in (False, True, False, True):
    h = pinned_empty(shape, torch.float32, write_combined=wc); h.copy_(src)
    assert h.is_pinned()
    ms = timed(h)
    ok = torch.equal(dev.cpu(), src)
    if rank == 0:
        gb = src.numel() * 4 / 1e9
        print(f"ranks={world} write_combined={wc}: {ms:.3f} ms/copy of {gb*1e3:.1f} MB per rank = {gb/ms*1e3:.1f} GB/s per GPU, "
              f"{gb*world/ms*1e3:.1f} GB/s aggregate, data ok={ok}", flush=True)
    del h
if world > 1: dist.destroy_process_group()
