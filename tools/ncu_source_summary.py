#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv`: executed warp-instructions per SASS opcode and (with --listing)
the SASS listing with per-instruction executed counts and stall samples. Usage: ncu_source_summary.py file.csv [--listing] [--kernel N]"""
import csv, sys, collections

def kernels(path):
    out, cur = [], None
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "Kernel Name":
                cur = {"name": row[1], "hdr": None, "rows": []}
                out.append(cur)
            elif cur is not None and cur["hdr"] is None:
                cur["hdr"] = row
            elif cur is not None:
                cur["rows"].append(row)
    return out

def main():
    path = sys.argv[1]
    listing = "--listing" in sys.argv
    which = int(sys.argv[sys.argv.index("--kernel") + 1]) if "--kernel" in sys.argv else 0
    k = kernels(path)[which]
    h = k["hdr"]
    i_src, i_ex, i_samp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[i_ex]) for r in k["rows"])
    print("kernel:", k["name"], " warp-instructions executed:", tot)
    by = collections.Counter()
    for r in k["rows"]:
        s = r[i_src].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        op = s.split()[0].split(".")[0]
        by[op] += int(r[i_ex])
    for op, n in by.most_common(40):
        print(f"  {op:12s} {n:12d} {100.0 * n / tot:6.2f} %")
    if listing:
        for n, r in enumerate(k["rows"]):
            print(f"{n:5d} {int(r[i_ex]):10d} {int(r[i_samp]):6d}  {r[i_src].strip()}")

if __name__ == "__main__":
    main()
