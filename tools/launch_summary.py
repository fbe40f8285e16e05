"""Summarise an ncu gpu__time_duration launch list (csv): time per kernel family and per (kernel, grid)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
tot, cnt, by_grid = collections.defaultdict(float), collections.Counter(), collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1e6 if row['Metric Unit'] == 'ns' else (v / 1e3 if row['Metric Unit'] == 'us' else v)
    base = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')
    tot[base] += v
    cnt[base] += 1
    g = by_grid[(base, row['Grid Size'])]
    g[0] += 1
    g[1] += v
T = sum(tot.values())
print(f"total {T:.2f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v:9.2f} ms {100 * v / T:5.1f}% n={cnt[k]:3d} {k[:60]}")
if len(sys.argv) > 2:
    print("--- per (kernel, grid) ---")
    for (k, g), (n, v) in sorted(by_grid.items(), key=lambda x: -x[1][1])[:int(sys.argv[2])]:
        print(f"{v:9.2f} ms n={n:3d} avg {v / n:7.3f} ms  {k[:40]:40s} {g}")
