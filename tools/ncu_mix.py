"""Instruction mix (warp-instructions per opcode, optionally per tile) from an `ncu --page source --csv` export.
Usage: python tools/ncu_mix.py file.csv [tiles]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ci = {n: i for i, n in enumerate(rows[hi])}
ops, tot = collections.Counter(), 0
for r in rows[hi + 1:]:
    n = int(r[ci["Instructions Executed"]] or 0)
    t = r[ci["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += n
    tot += n
print(f"total warp-instructions {tot}  ({tot / tiles:.0f} per tile)")
for k, v in ops.most_common(32):
    print(f"{k:10s} {v:13d} {v / tot:6.1%}  per tile {v / tiles:9.1f}")
