#!/bin/bash
# the other bench.py configurations (BASELINE.json configs 0/2/3/4) on the current build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 1 2 3; do
  python bench.py --config $c --steps 5 --warmup 3 --no-library-bar --no-cpu-baseline > gpurun_out/cfg$c.json 2> gpurun_out/cfg$c.err
  echo "config $c rc=$?"
done
python bench.py --config 4 --steps 2 > gpurun_out/cfg4.json 2> gpurun_out/cfg4.err
echo "config 4 rc=$?"
python - <<'PY'
import json
for c in (1, 2, 3, 4):
    try:
        d = json.loads(open(f"gpurun_out/cfg{c}.json").read().strip().splitlines()[-1])
        extra = [(s["global_batch"], round(s["value"])) for s in d["sweep"]] if "sweep" in d else d.get("parity_check", {}).get("ok")
        print("config", c, round(d["value"], 1), round(d["ms_per_step"], 2), d["config"].get("workload", "")[:70], extra)
    except Exception as e:
        print("config", c, "failed:", e)
PY
