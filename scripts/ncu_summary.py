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

if __name__ == "__main__":
    for rep in sys.argv[1:]:
        print("==", rep)
        for k in raw(rep):
            for key in KEYS:
                if key in k:
                    print(f"  {key}: {k[key][1]} {k[key][0]}")
        print("  top warp-stall reasons (% of samples):", stalls(rep))
