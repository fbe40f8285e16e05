"""Segment an `ncu --page source --csv --print-source sass` dump of ONE kernel at its BAR.SYNC instructions and print,
per segment (= phase between two CTA barriers, in program order): warp instructions executed, stall samples, and the
dominant opcodes.  Usage: python tools/ncu_phase_hist.py dump.csv [kernel-index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kern = -1
segs = [[0, 0, collections.Counter(), ""]]
for r in rows:
    if r and r[0] == "Kernel Name":
        kern += 1
        continue
    if kern != want or len(r) < 8 or not r[0].startswith("0x"):
        continue
    op = r[1].split()
    op = [t for t in op if not t.startswith("@")][0] if op else "?"
    try:
        n, smp = int(r[5]), int(r[4])
    except ValueError:
        continue
    s = segs[-1]
    s[0] += n
    s[1] += smp
    s[2][op.split(".")[0]] += n
    if op.startswith("BAR") or op.startswith("EXIT"):
        s[3] = op
        segs.append([0, 0, collections.Counter(), ""])
tot = sum(s[0] for s in segs)
ts = sum(s[1] for s in segs)
print(f"total warp-instructions {tot}, samples {ts}")
for i, (n, smp, ops, end) in enumerate(segs):
    if n == 0:
        continue
    top = " ".join(f"{k}:{100 * v / n:.0f}%" for k, v in ops.most_common(7))
    print(f"seg {i:2d} {100 * n / tot:5.1f}% inst {100 * smp / max(ts, 1):5.1f}% smp  | {top}")
