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
             read-back of the last image's maps, all inside the timed region
  roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs MEASURED_PEAKS hbm_gbs
  cpu_baseline  the reference's PyTorch-CPU op sequence (oracle/aten_port.py — /root/reference cannot
             travel to the GPU box) timed on the host cores on a bounded sample of the same workload
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
def cpu_reference_rate(sample_images, reps, threads):
    from oracle import aten_port            # test infrastructure: allowed here as the timed CPU baseline only
    torch.set_num_threads(threads)
    blobs, feats = synthetic(sample_images, M_BLOBS, CHANNELS, seed=0)
    run = lambda: aten_port.render(**blobs, features=feats, score_size=SIZE, interp_size=SIZE, ret_layout=False)
    run()                                    # warm-up
    times = []
    for _ in range(reps):
        t0 = time.perf_counter(); run(); times.append(time.perf_counter() - t0)
    pxb = sample_images * M_BLOBS * SIZE * SIZE
    return pxb / min(times) / 1e6, pxb / statistics.median(times) / 1e6, times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = 32
    blobs, feats = synthetic(sample, M_BLOBS, CHANNELS, seed=0)
    from oracle import aten_port
    torch.set_num_threads(threads)
    run = lambda: aten_port.render(**blobs, features=feats, score_size=SIZE, interp_size=SIZE, ret_layout=False)
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
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference's PyTorch-CPU op sequence (oracle/aten_port.py; bit-identical to /root/reference on the "
                "golden fixtures) on the host cores; the reference tree itself cannot travel to the GPU box",
    }))
    return 0


def workload_config(n_gpus):
    return {"workload": f"cfg5b: {N_IMG} images/GPU x {M_BLOBS} blobs, {SIZE}x{SIZE}, C={CHANNELS} feature grid, fp32 "
                        f"(BASELINE.json configs[4]; largest single-GPU config)",
            "images_per_gpu": N_IMG, "blobs": M_BLOBS, "size": SIZE, "channels": CHANNELS, "sharding": f"by-image x{n_gpus}",
            "l2": "outputs 6.5 GB/step stream through the 126 MB L2 (>> L2); the 87 MB of inputs are re-read each step",
            "e2e_pipeline": "HostRenderer: pinned H2D in 4 chunks on a copy stream, double-buffered staging (the copies of "
                            "step i+1 overlap the renders of step i), D2H of the last image's maps every step"}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
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

    import blobctrl_b200 as B
    from blobctrl_b200 import ops

    # this rank's image shard (weak scaling: N_IMG images per GPU, seeded per rank)
    host_blobs, host_feats = synthetic(N_IMG, M_BLOBS, CHANNELS, seed=rank)
    pin = {k: v.pin_memory() for k, v in host_blobs.items()}
    pin_feats = host_feats.pin_memory()
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

    t = torch.tensor([total, e2e_total], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total, e2e_total = t.tolist()
    pxb_step = N_IMG * M_BLOBS * P * world
    value = pxb_step * args.steps / total / 1e6
    e2e_value = pxb_step * args.steps / e2e_total / 1e6
    h2d = sum(v.numel() * v.element_size() for v in pin.values()) + pin_feats.numel() * 4
    d2h = out_host.numel() * 4

    peak, peak_src = peaks()
    dom = max(kern, key=lambda k: k["ms"])
    roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9, "peak": peak,
            "unit": "GB/s", "frac": dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9 / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this shape, from the
            # ncu --set full capture summarised in profiles/render_tc_r1.md (0.0872 GB read + 6.4000 GB written)
            "traffic": 6486408992 if dom["name"].startswith("render_tc") else None,
            "peak_source": peak_src, "alg_bytes_per_launch": dom["alg_bytes"], "avg_launch_ms": dom["ms"],
            "step_alg_bytes": algorithmic_bytes(N_IMG, M_BLOBS, P, CHANNELS, 4, 4),
            "step_frac": algorithmic_bytes(N_IMG, M_BLOBS, P, CHANNELS, 4, 4) / (total / args.steps) / 1e9 / peak,
            "kernels": kern}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_total / args.steps * 1e3},
            "gpu_launches": args.steps * len(kern), "roofline": roof, "clocks": clocks}

    if rank == 0 and world == 1:
        if not args.no_variants:
            line["variants"] = variants(B, ops, dev, peak)
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            best, med, times = cpu_reference_rate(256, 14, threads)
            line["cpu_baseline"] = {"value": best, "unit": UNIT, "cores": threads, "kind": "port", "median": med,
                                    "sample": f"256 of {N_IMG} images (x{M_BLOBS} blobs, {SIZE}x{SIZE}, C={CHANNELS}, fp32), "
                                              f"best of 14 after 1 warm-up, {sum(times):.1f} s CPU wall"}
            b1, _, t1 = cpu_reference_rate(8, 3, 1)
            line["cpu_baseline"]["single_thread"] = {"value": b1, "cores": 1, "sample": "8 images, best of 3"}
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


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
    ms = timed(lambda: B.splat_features_multiscale(**blobs, score_size=64, level_features=lf, out_dtype=torch.bfloat16))
    by = 64 * (28 * 32 + sum(33 * c * 2 + 33 * s * s * 2 + c * s * s * 2 for s, c in chans.items()))
    out["cfg3_multiscale_bf16"] = {"ms": ms, "Mpxblob_s": 64 * 32 * sum(s * s for s in chans) / ms / 1e3,
                                   "GBs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
    # latency-bound configs: report microseconds
    hb, hf = synthetic(1, 16, 320, seed=0)
    b2 = {k: v.to(dev) for k, v in hb.items()}; f2 = hf.to(dev)
    out["cfg2_latency_us"] = 1e3 * timed(lambda: B.splat_features(**b2, features=f2, score_size=64, interp_size=64,
                                                                   ret_layout=False), reps=50)
    hb, _ = synthetic(1, 1, 1, seed=0)
    b1 = {k: v.to(dev) for k, v in hb.items()}
    out["cfg1_dscore_512_latency_us"] = 1e3 * timed(lambda: B.splat_features(**b1, score_size=(512, 512),
                                                                              return_d_score=True), reps=50)
    out["cfg1_viz_512_latency_us"] = 1e3 * timed(lambda: B.splat_features(
        **b1, interp_size=64, viz_size=(512, 512), is_viz=True, score_size=64, viz_score_fn=B.viz_score_fn,
        viz_colors=B.BLOB_VIS_COLORS, only_vis=True), reps=50)
    # same preview as a CUDA graph over static buffers (host parameters in, one copy + one replay per call)
    from blobctrl_b200.preview import preview_renderer
    pr = preview_renderer((512, 512), dev)
    hx, hy, hc = hb["xs"].numpy(), hb["ys"].numpy(), hb["covs"].numpy()
    out["cfg1_viz_512_graph_latency_us"] = 1e3 * timed(lambda: pr(hx, hy, hc), reps=50)
    return out


if __name__ == "__main__":
    sys.exit(main())
