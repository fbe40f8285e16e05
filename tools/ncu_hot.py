"""Top SASS instructions by stall samples from an `ncu --page source --csv` export, with the dominant stall reasons and a
window of neighbouring instructions.  Usage: python tools/ncu_hot.py file.csv [--top 25] [--ctx 2]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
ctx = int(sys.argv[sys.argv.index("--ctx") + 1]) if "--ctx" in sys.argv else 0
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = rows[hdr_i + 1:]
ci = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot}, warp instructions {tot_inst}")
agg = {}
for n in stalls:
    agg[n] = sum(int(r[ci[n]] or 0) for r in data)
print("stall totals:", ", ".join(f"{k[6:]}={v / max(tot, 1):.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:top]
for i in sorted(order):
    for j in range(max(0, i - ctx), min(len(data), i + ctx + 1)):
        r = data[j]
        s = int(r[ci["# Samples"]] or 0)
        why = sorted(((int(r[ci[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
        mark = ">>" if j == i else "  "
        print(f"{mark}{j:5d} {s / max(tot, 1):6.2%} inst={int(r[ci['Instructions Executed']] or 0):9d} {r[ci['Source']].strip()[:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
    if ctx:
        print()
