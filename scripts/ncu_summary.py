"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io, collections

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        res.append({h: (u, v) for h, u, v in zip(hdr, units, vals)})
    return res

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in data:
        for h in cols:
            tot[h] += int(r[hdr.index(h)])
    s = sum(tot.values())
    return [(k, round(100 * v / s, 1)) for k, v in tot.most_common(6)]

if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "--traffic"):
    for rep in sys.argv[1:]:
        print("==", rep)
        for k in raw(rep):
            for key in KEYS:
                if key in k:
                    print(f"  {key}: {k[key][1]} {k[key][0]}")
        print("  top warp-stall reasons (% of samples):", stalls(rep))


def write_traffic(rep, kernel_prefix, shape, out_path, capture):
    """profiles/traffic.json entry for bench.py's roofline.traffic: dram bytes of ONE launch of the dominant kernel at the
    bench shape, tied to the kernel sources' hash (bench.kernel_source_sha) so a stale figure is never reported."""
    import json, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import kernel_source_sha
    rows = [k for k in raw(rep) if kernel_prefix in k["Kernel Name"][1]]
    if not rows:
        raise SystemExit(f"no launch of {kernel_prefix} in {rep}")
    k = rows[-1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(k["dram__bytes_read.sum"][1].replace(",", "")) * scale[k["dram__bytes_read.sum"][0]]
    wr = float(k["dram__bytes_write.sum"][1].replace(",", "")) * scale[k["dram__bytes_write.sum"][0]]
    entry = {"kernel_prefix": "render_tc", "kernel": k["Kernel Name"][1][:120], "shape": shape, "dram_bytes_read": rd,
             "dram_bytes_write": wr, "gpu_time_us": k["gpu__time_duration.sum"][1], "csrc_sha16": kernel_source_sha(),
             "capture": capture}
    try:
        doc = json.load(open(out_path))
    except Exception:
        doc = {"entries": []}
    doc["entries"] = [e for e in doc["entries"] if e.get("shape") != shape] + [entry]
    json.dump(doc, open(out_path, "w"), indent=1)
    print("wrote", out_path, entry)


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[1] == "--traffic":
    # python scripts/ncu_summary.py --traffic <rep> <capture label>   (cfg5b shape: the bench's dominant launch)
    write_traffic(sys.argv[2], "render_tc_kernel", {"N": 1024, "M": 64, "S": 64, "C": 320, "dtype": "f32"},
                  "profiles/traffic.json", sys.argv[3] if len(sys.argv) > 3 else "ncu --set full")
