"""Condense `ncu -i X.ncu-rep --page source --csv` (SASS view with sampling columns) into: stall-reason totals, opcode
mix, and the most-sampled instructions with their position in the kernel.

    ncu --set full --import-source on -k regex:<kernel> ... -o gpurun_out/x ; ncu -i gpurun_out/x.ncu-rep --page source --csv > x.csv
    python scripts/ncu_source_hotspots.py x.csv [n_top]
"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
    tot = sum(int(r[col["Instructions Executed"]]) for r in data)
    ts = sum(int(r[col["# Samples"]]) for r in data)
    print(f"{rows[0][1][:100] if len(rows[0]) > 1 else ''}\ntotal warp-instructions {tot}, static {len(data)}, samples {ts}")
    stalls = [c for c in hdr if c.startswith("stall_") and "(" not in c]
    st = collections.Counter()
    for r in data:
        for s in stalls:
            st[s] += int(r[col[s]] or 0)
    print("stall reasons:", ", ".join(f"{k[6:]} {v}" for k, v in st.most_common(8)))
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip()).split()[0]
        ops[op] += int(r[col["Instructions Executed"]])
        samp[op] += int(r[col["# Samples"]])
    print("opcode mix:", ", ".join(f"{op} {100 * n / tot:.1f}%" for op, n in ops.most_common(14)))
    top = sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]]))[:n_top]
    print("most-sampled instructions (index, samples, executed, long_sb / short_sb / barrier / wait, SASS):")
    for i, r in sorted(top):
        print(f"  {i:5d} {r[col['# Samples']]:>5s} {r[col['Instructions Executed']]:>8s}  {r[col['stall_long_sb']]:>4s}/{r[col['stall_short_sb']]:>4s}/"
              f"{r[col['stall_barrier']]:>4s}/{r[col['stall_wait']]:>4s}  {r[col['Source']].strip()[:80]}")


if __name__ == "__main__":
    main()
