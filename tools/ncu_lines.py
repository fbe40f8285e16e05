"""Per CUDA-source-line summary of an .ncu-rep (needs -lineinfo and the sources at the recorded paths):
top lines by stall samples and by executed warp-instructions, with the two dominant stall reasons.
Usage: python tools/ncu_lines.py report.ncu-rep [--top 40]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, lines = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        stall_cols = [(n, i) for i, n in enumerate(r) if n.startswith("stall_") and "Not Issued" not in n]
    elif hdr and r[0].strip().isdigit():
        num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
        s = num(r[hdr["# Samples"]])
        n = num(r[hdr["Instructions Executed"]])
        why = sorted(((num(r[i]), nm[6:]) for nm, i in stall_cols), reverse=True)[:2]
        lines.append((cur_file, int(r[0]), r[1].strip(), s, n, why))
ts, ti = sum(l[3] for l in lines), sum(l[4] for l in lines)
print(f"{rep}: {ts} samples, {ti} warp-instructions over {len(lines)} source lines")
print("--- by stall samples")
for f, ln, src, s, n, why in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{s / ts:6.2%} smp {n / ti:6.2%} inst  {f}:{ln:<5d} {src[:95]:95s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
print("--- by instructions")
for f, ln, src, s, n, why in sorted(lines, key=lambda l: -l[4])[:top // 2]:
    print(f"{s / ts:6.2%} smp {n / ti:6.2%} inst  {f}:{ln:<5d} {src[:95]}")
