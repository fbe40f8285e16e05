"""profiles/r2_traffic.json + launch-list summary from ONE ncu pass over a bench step:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:swn -s <launches of the warm-up steps> -c <launches of one step> --csv --log-file gpurun_out/step.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar --no-parity-check

    python tools/ncu_traffic.py gpurun_out/step.csv            # writes profiles/r2_traffic.json, prints the summary

`dram_bytes_per_step` per kernel family is what bench.py reports as `roofline.traffic` (divided by the launches per step).
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILY = [("swin_warp_block_kernel", "fused"), ("swin_fused_kernel", "fused"), ("swin_attn_stream_kernel", "fused"), ("swin_block_small", "fused"),
          ("mlp_persist_kernel", "mlp"), ("mlp_kernel", "mlp"), ("rowgemm", "rowgemm"), ("expand_warp", "rowgemm"), ("window_attn", "window_attn"),
          ("cross_attn", "cross_attn"), ("patch_embed", "heads"), ("conv_head", "heads"), ("bilinear", "heads")]


def family(name):
    for key, fam in FAMILY:
        if key in name:
            return fam
    return "glue"


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = {n: i for i, n in enumerate(rows[hi])}
    per = collections.defaultdict(lambda: collections.defaultdict(float))   # (id) -> metric -> value
    names = {}
    for r in rows[hi + 1:]:
        if len(r) <= hdr["Metric Value"]:
            continue
        kid = r[hdr["ID"]]
        names[kid] = r[hdr["Kernel Name"]]
        val = float(r[hdr["Metric Value"]].replace(",", ""))
        unit = r[hdr["Metric Unit"]]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(unit, 1.0)
        per[kid][r[hdr["Metric Name"]]] += val * scale
    fam = collections.defaultdict(lambda: {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
    kern = collections.defaultdict(lambda: {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
    for kid, m in per.items():
        for d, key in ((fam, family(names[kid])), (kern, names[kid].split("(")[0][:70])):
            d[key]["launches"] += 1
            d[key]["ms"] += m.get("gpu__time_duration.sum", 0.0)
            d[key]["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    tot = sum(v["ms"] for v in fam.values())
    print(f"{len(per)} launches, {tot:.2f} ms of kernel time (cold-cache, serialised: compare shares)")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"  {k:12s} {v['launches']:4d} launches {v['ms']:8.2f} ms {v['ms'] / tot:6.1%}  dram {v['dram_bytes'] / 1e9:7.2f} GB")
    print("per kernel:")
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"  {v['ms']:8.2f} ms {v['ms'] / tot:6.1%} n={v['launches']:3d} dram {v['dram_bytes'] / 1e9:6.2f} GB  {k}")
    out = {k: {"dram_bytes_per_step": v["dram_bytes"], "launches_per_step": v["launches"], "ncu_ms_per_step": v["ms"],
               "source": os.path.basename(path) + " (ncu, one bench step at batch 64)"} for k, v in fam.items()}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1])
