"""Per-opcode and per-region view of an .ncu-rep's SASS source page (read here, no GPU needed).
    python scripts/ncu_source.py REPORT TILES [marker-regex ...]
TILES = 128-pixel tiles the launch processed (to print warp instructions per tile).  Regions are the address ranges
between the first occurrences of the marker regexes (matched against the SASS text, in order)."""
import csv, io, re, subprocess, sys, collections

def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    return hdr, rows[2:]

def main():
    rep, tiles = sys.argv[1], float(sys.argv[2])
    markers = sys.argv[3:]
    hdr, data = load(rep)
    iS, iN, iX = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ops = collections.Counter()
    for r in data:
        op = r[iS].split()[0] if not r[iS].strip().startswith("@") else r[iS].split()[1]
        ops[op.split(".")[0]] += int(r[iX])
    total = sum(ops.values())
    print(f"warp instructions: {total}  = {total / tiles:.0f} per tile")
    print("  " + "  ".join(f"{k} {v / tiles:.0f}" for k, v in ops.most_common(24)))
    # regions
    bounds, pos = [0], 0
    for m in markers:
        rx = re.compile(m)
        for j in range(pos, len(data)):
            if rx.search(data[j][iS]):
                bounds.append(j); pos = j + 1
                break
        else:
            print("marker not found:", m)
    bounds.append(len(data))
    tot_samples = sum(int(r[iN]) for r in data) or 1
    for b in range(len(bounds) - 1):
        seg = data[bounds[b]:bounds[b + 1]]
        ns = sum(int(r[iN]) for r in seg)
        nx = sum(int(r[iX]) for r in seg)
        st = collections.Counter()
        for r in seg:
            for h, i in stall_cols:
                st[h[6:]] += int(r[i])
        s = sum(st.values()) or 1
        top = ", ".join(f"{k} {100 * v / s:.0f}%" for k, v in st.most_common(4))
        name = "start" if b == 0 else markers[b - 1]
        print(f"  [{bounds[b]:5d}..{bounds[b + 1]:5d}) from '{name}': samples {100 * ns / tot_samples:5.1f}%  instr/tile {nx / tiles:7.0f}  | {top}")

if __name__ == "__main__":
    main()
