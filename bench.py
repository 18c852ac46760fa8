#!/usr/bin/env python
"""bench.py — blob-splat throughput on B200 (BASELINE.json metric: Mpixel*blob/s + % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config 5 — batch-sharded blob rendering, 1024 images x 64 blobs at
64x64 latent resolution with the C=320 feature grid, float32 (SURVEY.md §8(d) variant 5b), PER GPU
(weak scaling: images shard by rank, no collective on the hot path).  A "step" is one full render of
that batch through the public API: blob parameters + features -> composed score maps [N,65,64,64] and
feature grid [N,320,64,64].

  value      whole-job Mpixel*blob/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public host-input API (blobctrl_b200.streaming.HostRenderer) with HOST
             (pinned) inputs: chunked H2D of parameters and features overlapped with the render, and a D2H
             read-back of the last image's maps (a SAMPLE of the result: the maps are consumed on the device by
             BlobNet/UNet), all inside the timed region.  e2e.full_d2h is the same step with the WHOLE 6.46 GB
             result copied back, measured once beside it; e2e.h2d_probe is the raw aggregated pinned-H2D rate of the
             same bytes (the host link's ceiling at this N).
  roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs MEASURED_PEAKS hbm_gbs;
             roofline.sustained is the same step back to back for >= 2 s (settled clocks), with its own frac
  cpu_baseline  the UNMODIFIED reference renderer (baseline/_ref, installed by scripts/install_reference.sh; kind
             "reference") — or, when that is absent, the reference's PyTorch-CPU op sequence restated in
             oracle/aten_port.py (kind "port") — timed on the host cores on a bounded sample of the same workload
  strong / gather (N > 1)  BASELINE configs[4] literally: 1024 images in total, 1024/N per rank; and the optional
             final NCCL all-gather of the maps (sharding.gather_maps), timed separately, never inside `value`
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IMG, M_BLOBS, SIZE, CHANNELS = 1024, 64, 64, 320
METRIC = "blob_splat_throughput"
UNIT = "Mpixel*blob/s"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synthetic(n, m, c, seed, device="cpu", dtype=torch.float32):
    """SURVEY.md §8(d) input distribution (same as oracle.blob_oracle.synthetic_blobs, torch RNG)."""
    g = torch.Generator().manual_seed(seed)
    xs, ys = torch.rand(n, m, generator=g), torch.rand(n, m, generator=g)
    a = 0.02 + 0.2 * torch.rand(n, m, generator=g)
    b = 0.02 + 0.2 * torch.rand(n, m, generator=g)
    th = torch.pi * torch.rand(n, m, generator=g)
    cs, sn = torch.cos(th), torch.sin(th)
    rot = torch.stack([cs, sn, -sn, cs], -1).view(n, m, 2, 2)
    covs = rot @ torch.diag_embed(torch.stack([a * a, b * b], -1)) @ rot.transpose(-1, -2)
    sizes = (torch.rand(n, m, generator=g) > 0.1).float()
    feats = torch.randn(n, m + 1, c, generator=g)
    return {"xs": xs.to(dtype), "ys": ys.to(dtype), "covs": covs.to(dtype), "sizes": sizes}, feats.to(dtype)


def algorithmic_bytes(n, m, p, c, e_f, e_o, scores=True, grid=True):
    """params (7 fp32/blob) + features read + score maps written + feature grid written."""
    k = m + 1
    return n * (28 * m + (k * c * e_f if grid else 0) + (k * p * e_o if scores else 0) + (c * p * e_o if grid else 0))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference's PyTorch-CPU renderer (ATen port), bounded sample
# ----------------------------------------------------------------------------------------------------
def reference_renderer():
    """(run(blobs, feats), kind): the unmodified reference from baseline/_ref when installed, else the ATen port."""
    try:
        from baseline import ref_loader
        if ref_loader.available():
            U = ref_loader.load(need_pipeline=False).utils
            return (lambda blobs, feats: U.splat_features(**blobs, features=feats, score_size=SIZE, interp_size=SIZE,
                                                          ret_layout=False)), "reference"
    except Exception as e:  # pragma: no cover
        print(f"bench: baseline/_ref unusable ({e}); falling back to the ATen port", file=sys.stderr)
    from oracle import aten_port            # test infrastructure: allowed here as the timed CPU baseline only
    return (lambda blobs, feats: aten_port.render(**blobs, features=feats, score_size=SIZE, interp_size=SIZE,
                                                  ret_layout=False)), "port"


def cpu_reference_rate(sample_images, reps, threads):
    render, kind = reference_renderer()
    torch.set_num_threads(threads)
    blobs, feats = synthetic(sample_images, M_BLOBS, CHANNELS, seed=0)
    run = lambda: render(blobs, feats)
    run()                                    # warm-up
    times = []
    for _ in range(reps):
        t0 = time.perf_counter(); run(); times.append(time.perf_counter() - t0)
    pxb = sample_images * M_BLOBS * SIZE * SIZE
    return pxb / min(times) / 1e6, pxb / statistics.median(times) / 1e6, times, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = 32
    blobs, feats = synthetic(sample, M_BLOBS, CHANNELS, seed=0)
    render, kind = reference_renderer()
    torch.set_num_threads(threads)
    run = lambda: render(blobs, feats)
    for _ in range(max(args.warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    pxb = sample * M_BLOBS * SIZE * SIZE * args.steps
    val = pxb / dt / 1e6
    desc = f"{sample} images x {M_BLOBS} blobs x {SIZE}x{SIZE} x C={CHANNELS} fp32 per step (1/32 of the per-GPU batch)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("the UNMODIFIED reference renderer (blobctrl/utils/utils.py::splat_features, pip-installed into baseline/_ref) "
                 "on the host cores" if kind == "reference" else
                 "reference's PyTorch-CPU op sequence (oracle/aten_port.py; bit-identical to /root/reference on the golden "
                 "fixtures) on the host cores; baseline/_ref was not installed"),
    }))
    return 0


def workload_config(n_gpus):
    return {"workload": f"cfg5b: {N_IMG} images/GPU x {M_BLOBS} blobs, {SIZE}x{SIZE}, C={CHANNELS} feature grid, fp32 "
                        f"(BASELINE.json configs[4]; largest single-GPU config)",
            "images_per_gpu": N_IMG, "blobs": M_BLOBS, "size": SIZE, "channels": CHANNELS, "sharding": f"by-image x{n_gpus}",
            "l2": "outputs 6.5 GB/step stream through the 126 MB L2 (>> L2); the 87 MB of inputs are re-read each step",
            "e2e_pipeline": "H2D + render + sample D2H — HostRenderer: pinned H2D on a copy stream (parameters whole, features in 4 chunks), double-buffered "
                            "staging (the copies of step i+1 overlap the renders of step i), D2H of the LAST IMAGE's maps every "
                            "step (6.3 MB of the 6.46 GB result; the maps are consumed on the device).  e2e.full_d2h copies the "
                            "whole result back instead"}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def time_steps(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(); barrier()
    return e0.elapsed_time(e1) / 1e3   # seconds for `steps` steps


def kernel_source_sha():
    """Hash of the sources the dominant kernel is built from: ties a committed ncu traffic figure to the running build."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "blobctrl_b200", "csrc")
    for f in ("common.cuh", "render_tc.cuh"):      # the float32 fused kernel lives in these two (render_tc2.cuh: 16-bit only)
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full summary
    (profiles/traffic.json, written by scripts/ncu_summary.py) — only when it was captured on THIS build of the kernel
    at THIS shape; otherwise null (a stale number is worse than none)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None, "no profiles/traffic.json"
    for e in t.get("entries", []):
        if (kernel_name.startswith(e.get("kernel_prefix", "?")) and e.get("csrc_sha16") == kernel_source_sha()
                and e.get("shape") == {"N": N_IMG, "M": M_BLOBS, "S": SIZE, "C": CHANNELS, "dtype": "f32"}):
            return int(e["dram_bytes_read"] + e["dram_bytes_write"]), f"profiles/traffic.json ({e.get('capture', '?')})"
    return None, "profiles/traffic.json has no capture of this build/shape"


def run_for(fn, seconds, barrier):
    """fn back to back for >= `seconds` of device time: (ms per call, calls).  Clocks settle under sustained load."""
    torch.cuda.synchronize(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    calls, batch = 0, 50
    e0.record()
    t0 = time.perf_counter()
    while True:
        for _ in range(batch):
            fn()
        calls += batch
        torch.cuda.synchronize()
        if time.perf_counter() - t0 >= seconds:
            break
    e1.record(); torch.cuda.synchronize(); barrier()
    return e0.elapsed_time(e1) / calls, calls


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: dist.barrier()) if dist else (lambda: None)

    def max_over_ranks(*vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    import blobctrl_b200 as B
    from blobctrl_b200 import ops

    # this rank's image shard (weak scaling: N_IMG images per GPU, seeded per rank)
    host_blobs, host_feats = synthetic(N_IMG, M_BLOBS, CHANNELS, seed=rank)
    # host inputs of the e2e path: page-locked; BLOBSPLAT_BENCH_WC=1 allocates them write-combined (blobctrl_b200.hostmem)
    wc_inputs = os.environ.get("BLOBSPLAT_BENCH_WC", "0") == "1"
    from blobctrl_b200.hostmem import pinned_like
    pin = pinned_like(host_blobs, write_combined=wc_inputs)
    pin_feats = pinned_like({"f": host_feats}, write_combined=wc_inputs)["f"]
    blobs = {k: v.to(dev) for k, v in host_blobs.items()}
    feats = host_feats.to(dev)
    P = SIZE * SIZE

    def step():
        return B.splat_features(**blobs, features=feats, score_size=SIZE, interp_size=SIZE, ret_layout=False)

    out = step(); torch.cuda.synchronize()
    del out

    # ---- per-kernel durations (CUDA events between the launches of one step) -----------------------
    with ClockSampler(local) as clk:
        kern = probe_kernels(ops, blobs, feats, dev, max(args.steps // 2, 5))
        total = time_steps(step, args.steps, args.warmup, barrier)
        # ---- e2e: host (pinned) inputs in, last image's maps out, inside the timed region ----------
        out_host = torch.empty((M_BLOBS + 1 + CHANNELS, SIZE, SIZE), dtype=torch.float32).pin_memory()

        from blobctrl_b200.streaming import HostRenderer
        host_renderer = HostRenderer(N_IMG, M_BLOBS, SIZE, CHANNELS, torch.float32, dev, chunks=4)

        def e2e_step():
            # public host-input API: chunked H2D on a copy stream overlapped with the fused render
            o = host_renderer(pin["xs"], pin["ys"], pin["covs"], pin["sizes"], pin_feats)
            out_host[:M_BLOBS + 1].copy_(o["scores_pyramid"][SIZE][-1], non_blocking=True)
            out_host[M_BLOBS + 1:].copy_(o["feature_grid"][-1], non_blocking=True)

        e2e_total = time_steps(e2e_step, args.steps, args.warmup, barrier)
    clocks = clk.summary()

    # ---- the host link's ceiling at this N: the same H2D bytes, all ranks at once, nothing else running -------------
    stage = {k: torch.empty_like(v, device=dev) for k, v in pin.items()}
    stage_f = torch.empty_like(pin_feats, device=dev)

    def h2d_only():
        for k, v in pin.items():
            stage[k].copy_(v, non_blocking=True)
        stage_f.copy_(pin_feats, non_blocking=True)
    h2d_total = time_steps(h2d_only, args.steps, 3, barrier)
    del stage, stage_f

    # ---- sustained: the same device-resident step back to back for >= 2 s (clocks settle, power state ramps) -------
    with ClockSampler(local) as sclk:
        sus_ms, sus_calls = run_for(step, args.sustained_seconds, barrier)
    sus_clocks = sclk.summary()

    total, e2e_total, h2d_total, sus_ms = max_over_ranks(total, e2e_total, h2d_total, sus_ms)
    pxb_step = N_IMG * M_BLOBS * P * world
    value = pxb_step * args.steps / total / 1e6
    e2e_value = pxb_step * args.steps / e2e_total / 1e6
    h2d = sum(v.numel() * v.element_size() for v in pin.values()) + pin_feats.numel() * 4
    d2h = out_host.numel() * 4

    peak, peak_src = peaks()
    step_bytes = algorithmic_bytes(N_IMG, M_BLOBS, P, CHANNELS, 4, 4)
    dom = max(kern, key=lambda k: k["ms"])
    traffic, traffic_src = measured_traffic(dom["name"])
    # a step of this workload IS one launch of the dominant kernel (gpu_launches = steps): its average launch duration over the
    # timed region is the region's CUDA-event time / steps.  The same launch bracketed by its own pair of events, in a loop
    # of its own before the timed region, is kept beside it (events between launches cost ~2 % at this size).
    launch_ms = total / args.steps * 1e3 if len(kern) == 1 else dom["ms"]
    roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["alg_bytes"] / (launch_ms * 1e-3) / 1e9, "peak": peak,
            "unit": "GB/s", "frac": dom["alg_bytes"] / (launch_ms * 1e-3) / 1e9 / peak,
            "traffic": traffic, "traffic_source": traffic_src, "kernel_source_sha16": kernel_source_sha(),
            "peak_source": peak_src, "alg_bytes_per_launch": dom["alg_bytes"], "avg_launch_ms": launch_ms,
            "avg_launch_ms_bracketed": dom["ms"], "frac_bracketed": dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9 / peak,
            "step_alg_bytes": step_bytes, "step_frac": step_bytes / (total / args.steps) / 1e9 / peak,
            "sustained": {"seconds": args.sustained_seconds, "steps": sus_calls, "ms_per_step": sus_ms,
                          "achieved": step_bytes / (sus_ms * 1e-3) / 1e9, "frac": step_bytes / (sus_ms * 1e-3) / 1e9 / peak,
                          "value": pxb_step / (sus_ms * 1e-3) / 1e6, "clocks": sus_clocks,
                          "note": "same step as `value`, back to back for >= the stated seconds; max over ranks"},
            "kernels": kern}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_total / args.steps * 1e3,
                    "what": "H2D of all inputs + render + D2H of the last image's maps (a sample of the result)",
                    "host_inputs": "pinned, write-combined" if wc_inputs else "pinned",
                    "h2d_probe": {"ms_per_step": h2d_total / args.steps * 1e3,
                                  "aggregate_GBs": h2d * world / (h2d_total / args.steps) / 1e9,
                                  "per_gpu_GBs": h2d / (h2d_total / args.steps) / 1e9,
                                  "note": "raw pinned H2D of one step's inputs on every rank at once (max over ranks): the "
                                          "floor of an e2e step at this N; the box is one NUMA node, 32 vCPUs"}},
            "gpu_launches": args.steps * len(kern), "roofline": roof, "clocks": clocks}

    # ---- e2e with the WHOLE result copied back (once, few steps: 6.46 GB of pinned host memory) ---------------------
    if rank == 0 and world == 1 and not args.no_variants:
        try:
            full_s = torch.empty((N_IMG, M_BLOBS + 1, SIZE, SIZE), dtype=torch.float32).pin_memory()
            full_g = torch.empty((N_IMG, CHANNELS, SIZE, SIZE), dtype=torch.float32).pin_memory()

            def e2e_full():
                o = host_renderer(pin["xs"], pin["ys"], pin["covs"], pin["sizes"], pin_feats)
                full_s.copy_(o["scores_pyramid"][SIZE], non_blocking=True)
                full_g.copy_(o["feature_grid"], non_blocking=True)
            t_full = time_steps(e2e_full, 3, 1, barrier)
            nbytes = (full_s.numel() + full_g.numel()) * 4
            line["e2e"]["full_d2h"] = {"value": pxb_step * 3 / t_full / 1e6, "ms_per_step": t_full / 3 * 1e3,
                                       "d2h_bytes_per_step": nbytes, "d2h_GBs": nbytes / (t_full / 3) / 1e9, "steps": 3,
                                       "what": "the same step with the WHOLE result (score maps + feature grid) copied to pinned "
                                               "host memory: bounded by the D2H link, not by the render"}
            del full_s, full_g
        except Exception as e:  # pragma: no cover
            line["e2e"]["full_d2h"] = {"error": str(e)[:200]}

    # ---- N > 1: strong scaling (BASELINE configs[4] literally) and the optional final gather, timed separately -----
    if world > 1:
        from blobctrl_b200.sharding import gather_maps, shard_bounds
        lo, hi = shard_bounds(N_IMG, rank, world)
        sb = {k: v[: hi - lo].contiguous() for k, v in blobs.items()}
        sf = feats[: hi - lo].contiguous()

        def strong_step():
            return B.splat_features(**sb, features=sf, score_size=SIZE, interp_size=SIZE, ret_layout=False)
        t_strong = time_steps(strong_step, args.steps, args.warmup, barrier)
        with ClockSampler(local):
            strong_sus, _ = run_for(strong_step, 1.0, barrier)
        o = strong_step()
        maps = torch.cat([o["scores_pyramid"][SIZE], o["feature_grid"]], 1)      # [N/G, 65 + 320, 64, 64]
        del o
        gather_maps(maps, N_IMG); torch.cuda.synchronize()
        t_gather = time_steps(lambda: gather_maps(maps, N_IMG), 5, 2, barrier)
        t_strong, strong_sus, t_gather = max_over_ranks(t_strong, strong_sus, t_gather)
        gbytes = maps.numel() * 4 * world
        line["strong"] = {"images_total": N_IMG, "images_per_gpu": hi - lo, "ms_per_step": t_strong / args.steps * 1e3,
                          "value": N_IMG * M_BLOBS * P * args.steps / t_strong / 1e6,
                          "sustained_ms_per_step": strong_sus, "sustained_value": N_IMG * M_BLOBS * P / (strong_sus * 1e-3) / 1e6,
                          "note": "strong scaling: 1024 images in total; compare value with the N=1 line's value"}
        line["gather"] = {"ms": t_gather / 5 * 1e3, "bytes_total": gbytes,
                          "bus_GBs_per_gpu": gbytes * (world - 1) / world / (t_gather / 5) / 1e9,
                          "what": "sharding.gather_maps: NCCL all_gather_into_tensor of the strong-scaling shard's maps "
                                  "([1024/N, 385, 64, 64] fp32 per rank) to every rank; never inside `value`"}
        del maps

    if rank == 0 and world == 1:
        if not args.no_variants:
            line["variants"] = variants(B, ops, dev, peak)
            if not args.no_cfg4:
                line["variants"]["cfg4_edit_loop"] = cfg4_variant()
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            best, med, times, kind = cpu_reference_rate(256, 14, threads)
            line["cpu_baseline"] = {"value": best, "unit": UNIT, "cores": threads, "kind": kind, "median": med,
                                    "sample": f"256 of {N_IMG} images (x{M_BLOBS} blobs, {SIZE}x{SIZE}, C={CHANNELS}, fp32), "
                                              f"best of 14 after 1 warm-up, {sum(times):.1f} s CPU wall"}
            b1, _, t1, _ = cpu_reference_rate(8, 3, 1)
            line["cpu_baseline"]["single_thread"] = {"value": b1, "cores": 1, "sample": "8 images, best of 3"}
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


def cfg4_variant():
    """BASELINE configs[3]: the reference's 50-step edit loop (random-init SD-1.5 shapes, batch 8 -> 16 with CFG, fp16 +
    autocast) stock vs with the CUDA splat / N1 / N2 / N4 substituted (baseline/cfg4_harness.py).  s/edit per arm, final
    latent agreement, and the splat's measured share of an edit."""
    try:
        from baseline import cfg4_harness, ref_loader
        if not ref_loader.available():
            return {"unavailable": "baseline/_ref not installed (scripts/install_reference.sh needs /root/reference)"}
        r = cfg4_harness.compare_arms(device="cuda", dtype=torch.float16, batch=8, steps=50)
        torch.cuda.empty_cache()
        return r
    except Exception as e:  # pragma: no cover
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}


def probe_kernels(ops, blobs, feats, dev, reps):
    """Average device time of each kernel of one step, CUDA events on the launching stream."""
    P = SIZE * SIZE
    xs, ys, covs, sizes, n, m = ops.canonical_blobs(**blobs)
    kern = []
    try:
        ops.render_fused(xs, ys, covs, sizes, feats, SIZE, SIZE)
        fused = True
    except Exception:
        fused = False
    ev = lambda: torch.cuda.Event(enable_timing=True)
    if fused:
        ts = []
        for i in range(reps + 2):
            a, b = ev(), ev(); a.record(); ops.render_fused(xs, ys, covs, sizes, feats, SIZE, SIZE); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        kern.append({"name": "render_tc (stages 1+2+3 fused, tcgen05)", "ms": statistics.mean(ts[2:]),
                     "alg_bytes": algorithmic_bytes(n, m, P, CHANNELS, 4, 4)})
        return kern
    t1, t2 = [], []
    for i in range(reps + 2):
        a, b, c = ev(), ev(), ev()
        a.record(); d, _ = ops.render_scores(xs, ys, covs, sizes, SIZE, SIZE); b.record()
        g = ops.feature_splat(d, feats); c.record()
        torch.cuda.synchronize(); t1.append(a.elapsed_time(b)); t2.append(b.elapsed_time(c))
        del d, g
    kern.append({"name": "scores_lane_pixel_f32 (stages 1+2)", "ms": statistics.mean(t1[2:]),
                 "alg_bytes": algorithmic_bytes(n, m, P, CHANNELS, 4, 4, scores=True, grid=False)})
    kern.append({"name": "feature_splat_fma (stage 3)", "ms": statistics.mean(t2[2:]),
                 # reads the composed maps back (K*P*4) + features, writes the grid
                 "alg_bytes": n * ((m + 1) * CHANNELS * 4 + CHANNELS * P * 4)})
    return kern


def variants(B, ops, dev, peak):
    """Side measurements (few steps each): cfg5a scores only, cfg5c bf16, cfg3 multi-scale bf16,
    stage-2 mapping A/B, latency of the two small configs.  Reported, not the headline."""
    out = {}
    P = SIZE * SIZE

    def timed(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    hb, hf = synthetic(N_IMG, M_BLOBS, CHANNELS, seed=0)
    blobs = {k: v.to(dev) for k, v in hb.items()}
    feats = hf.to(dev)
    xs, ys, covs, sizes, n, m = ops.canonical_blobs(**blobs)
    for mode in ("lane_pixel", "warp_scan"):
        ms = timed(lambda: ops.render_scores(xs, ys, covs, sizes, SIZE, SIZE, composite_mode=mode))
        by = algorithmic_bytes(n, m, P, 0, 4, 4, grid=False)
        out[f"cfg5a_scores_fp32_{mode}"] = {"ms": ms, "Mpxblob_s": n * m * P / ms / 1e3, "GBs": by / ms / 1e6,
                                            "frac": by / ms / 1e6 / peak}
    out["cfg5a_scores_fp32_warp_scan"]["note"] = ("lane = blob suffix scan (BLOBSPLAT_COMPOSITE_WARP_SCAN): an explicit opt-in mode "
                                                  "kept for A/B; AUTO never selects it (lane = pixel is faster at every shape)")
    try:
        fb = feats.to(torch.bfloat16)
        ms = timed(lambda: B.splat_features(**blobs, features=fb, score_size=SIZE, interp_size=SIZE, ret_layout=False,
                                            out_dtype=torch.bfloat16))
        by = algorithmic_bytes(n, m, P, CHANNELS, 2, 2)
        out["cfg5c_bf16"] = {"ms": ms, "Mpxblob_s": n * m * P / ms / 1e3, "GBs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
    except Exception as e:  # pragma: no cover
        out["cfg5c_bf16"] = {"error": str(e)[:200]}
    del blobs, feats
    # cfg3: 32 blobs, N=64, levels 64/32/16/8 with C=320/640/1280/1280, bf16 maps
    hb, _ = synthetic(64, 32, 1, seed=0)
    blobs = {k: v.to(dev) for k, v in hb.items()}
    chans = {64: 320, 32: 640, 16: 1280, 8: 1280}
    g = torch.Generator().manual_seed(1)
    lf = {s: torch.randn(64, 33, c, generator=g).to(dev).to(torch.bfloat16) for s, c in chans.items()}
    ms = timed(lambda: B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16), reps=50, warm=5)
    by = 64 * (28 * 32 + sum(33 * c * 2 + 33 * s * s * 2 + c * s * s * 2 for s, c in chans.items()))
    out["cfg3_multiscale_bf16"] = {"ms": ms, "Mpxblob_s": 64 * 32 * sum(s * s for s in chans) / ms / 1e3,
                                   "GBs": by / ms / 1e6, "frac": by / ms / 1e6 / peak,
                                   "launches": ["render_tc2<kPyr> (stages 1+2+3 at 64x64 + pyramid levels 32/16/8)",
                                                "splat_tma (stage 3 at 32/16/8, TMA operands)"],
                                   "what": "ONE call (blobsplat_render_multiscale), eager; graph_ms = the same call as a CUDA graph"}
    # stage 3 alone at the cfg5c shape (1024 images, K = 65, C = 320, bf16) on the TMA engine vs the thread-staged tensor engine
    try:
        sc5 = torch.rand(N_IMG, M_BLOBS + 1, SIZE, SIZE, device=dev)
        sc5 = (sc5 / sc5.sum(1, keepdim=True)).to(torch.bfloat16)
        f5 = torch.randn(N_IMG, M_BLOBS + 1, CHANNELS, device=dev).to(torch.bfloat16)
        by5 = N_IMG * ((M_BLOBS + 1) * (CHANNELS + P) * 2 + CHANNELS * P * 2)
        s3 = {}
        for eng in ("tma", "tensor"):
            ms5 = timed(lambda: ops.feature_splat(sc5, f5, engine=eng))
            s3[eng] = {"ms": ms5, "GBs": by5 / ms5 / 1e6, "frac": by5 / ms5 / 1e6 / peak}
        out["cfg5c_stage3_only_bf16"] = s3
        del sc5, f5
    except Exception as e:  # pragma: no cover
        out["cfg5c_stage3_only_bf16"] = {"error": str(e)[:200]}
    # latency-bound configs: report microseconds.  eager = the reference-signature call; graph = the same call with
    # cuda_graph=True (one cudaGraphLaunch per call, blobctrl_b200/graphs.py); reference_* = the UNMODIFIED reference
    # function (baseline/_ref) on the host CPU as its scripts run it, and its ~20-launch ATen sequence on this GPU
    REF = None
    try:
        from baseline import ref_loader
        if ref_loader.available():
            REF = ref_loader.load(need_pipeline=False).utils
    except Exception:  # pragma: no cover
        REF = None

    def cpu_us(fn, reps=5):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e6

    lat = {}
    hb, hf = synthetic(1, 16, 320, seed=0)
    b2 = {k: v.to(dev) for k, v in hb.items()}; f2 = hf.to(dev)
    kw2 = dict(features=f2, score_size=64, interp_size=64, ret_layout=False)
    lat["cfg2"] = {"eager_us": 1e3 * timed(lambda: B.splat_features(**b2, **kw2), reps=50),
                   "graph_us": 1e3 * timed(lambda: B.splat_features(**b2, **kw2, cuda_graph=True), reps=200)}
    hb1, _ = synthetic(1, 1, 1, seed=0)
    b1 = {k: v.to(dev) for k, v in hb1.items()}
    kwd = dict(score_size=(512, 512), return_d_score=True)
    lat["cfg1_dscore_512"] = {"eager_us": 1e3 * timed(lambda: B.splat_features(**b1, **kwd), reps=50),
                              "graph_us": 1e3 * timed(lambda: B.splat_features(**b1, **kwd, cuda_graph=True), reps=200)}
    colors = B.BLOB_VIS_COLORS.to(dev)
    kwv = dict(interp_size=64, viz_size=(512, 512), is_viz=True, score_size=64, viz_score_fn=B.viz_score_fn,
               viz_colors=colors, only_vis=True)
    lat["cfg1_viz_512"] = {"eager_us": 1e3 * timed(lambda: B.splat_features(**b1, **kwv), reps=50),
                           "graph_us": 1e3 * timed(lambda: B.splat_features(**b1, **kwv, cuda_graph=True), reps=200)}
    # the app's call shape: host numpy parameters in (blobctrl_app.py:604-646), one H2D copy + one graph replay per call
    from blobctrl_b200.preview import preview_renderer
    pr = preview_renderer((512, 512), dev)
    hx, hy, hc = hb1["xs"].numpy(), hb1["ys"].numpy(), hb1["covs"].numpy()
    lat["cfg1_viz_512"]["graph_host_params_us"] = 1e3 * timed(lambda: pr(hx, hy, hc), reps=100)
    # the UI's whole preview step, synchronous, host to host (blobctrl_app.py:637-648 up to Image.fromarray): host parameters
    # in, host uint8 picture out — one graph replay (H2D of 28 bytes, preview launch writing bytes, pinned D2H of 786 KB)
    prp = preview_renderer((512, 512), dev, picture=True)
    for _ in range(5):
        prp.render_picture(hx, hy, hc)
    t0 = time.perf_counter()
    for _ in range(200):
        prp.render_picture(hx, hy, hc)
    lat["cfg1_viz_512"]["picture_host_to_host_us"] = (time.perf_counter() - t0) / 200 * 1e6
    if REF is not None:
        torch.set_num_threads(os.cpu_count() or 1)
        c64 = {k: (v.double() if k != "sizes" else v) for k, v in hb1.items()}            # the scripts feed float64
        kwv_cpu = dict(kwv, viz_colors=REF.BLOB_VIS_COLORS, viz_score_fn=REF.viz_score_fn)
        lat["cfg1_viz_512"]["reference_cpu_f64_us"] = cpu_us(lambda: REF.splat_features(**c64, **kwv_cpu))
        import numpy as _np
        lat["cfg1_viz_512"]["reference_cpu_f64_picture_us"] = cpu_us(
            lambda: (REF.splat_features(**c64, **kwv_cpu)["feature_img"][0].permute(1, 2, 0).contiguous().cpu().numpy() * 255).astype(_np.uint8))
        lat["cfg1_dscore_512"]["reference_cpu_f64_us"] = cpu_us(lambda: REF.splat_features(**c64, **kwd))
        lat["cfg2"]["reference_cpu_f32_us"] = cpu_us(lambda: REF.splat_features(**hb, features=hf, score_size=64, interp_size=64,
                                                                                 ret_layout=False))
        try:    # the same reference code with its tensors on this GPU (device-agnostic torch ops)
            kwv_gpu = dict(kwv, viz_colors=REF.BLOB_VIS_COLORS.to(dev), viz_score_fn=REF.viz_score_fn)
            lat["cfg1_viz_512"]["reference_on_gpu_us"] = 1e3 * timed(lambda: REF.splat_features(**b1, **kwv_gpu), reps=20)
            lat["cfg1_dscore_512"]["reference_on_gpu_us"] = 1e3 * timed(lambda: REF.splat_features(**b1, **kwd), reps=20)
            lat["cfg2"]["reference_on_gpu_us"] = 1e3 * timed(lambda: REF.splat_features(**b2, **kw2), reps=20)
            # and at the headline shape, 256-image chunks (its [N,M,2,P] intermediates are 0.5 GB each per chunk)
            hbb, hff = synthetic(256, M_BLOBS, CHANNELS, seed=0)
            bb = {k: v.to(dev) for k, v in hbb.items()}; ff = hff.to(dev)
            ms = timed(lambda: REF.splat_features(**bb, features=ff, score_size=SIZE, interp_size=SIZE, ret_layout=False),
                       reps=3, warm=1)
            out["cfg5b_reference_on_gpu"] = {"ms_per_256_images": ms, "Mpxblob_s": 256 * M_BLOBS * P / ms / 1e3,
                                             "what": "the unmodified reference splat_features with CUDA tensors on this B200 "
                                                     "(ATen/cuSOLVER/cuBLAS kernels), 256-image chunk"}
            del bb, ff
        except Exception as e:  # pragma: no cover
            lat["reference_on_gpu_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    out["latency"] = lat
    # cfg3 as a CUDA graph (one launch of the captured call sequence)
    try:
        g3 = torch.cuda.CUDAGraph()
        B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16); torch.cuda.synchronize()
        with torch.cuda.graph(g3):
            keep = B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16)
        ms = timed(g3.replay, reps=50)
        out["cfg3_multiscale_bf16"]["graph_ms"] = ms
        out["cfg3_multiscale_bf16"]["graph_frac"] = by / ms / 1e6 / peak
        del keep, g3
    except Exception as e:  # pragma: no cover
        out["cfg3_multiscale_bf16"]["graph_error"] = str(e)[:200]
    return out


if __name__ == "__main__":
    sys.exit(main())
