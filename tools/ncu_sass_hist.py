"""Opcode histogram (executed instructions and stall samples) from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [r for r in rows[1:] if r[ie].isdigit()]
tot = sum(int(r[ie]) for r in data)
tots = sum(int(r[isamp]) for r in data)
print("total warp-inst", tot, "samples", tots, "sass lines", len(data))
h, hs = collections.Counter(), collections.Counter()
for r in data:
    toks = r[ia].strip().split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    h[op] += int(r[ie])
    hs[op] += int(r[isamp])
for op, c in h.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:12s} {c / tot * 100:5.1f}% inst   {hs[op] / max(tots, 1) * 100:5.1f}% samples")
if len(sys.argv) > 3:
    print("--- top sampled SASS lines ---")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:int(sys.argv[3])]:
        print(f"{int(r[isamp]):7d} smp {int(r[ie]):10d} ex  {r[ia].strip()[:100]}")
