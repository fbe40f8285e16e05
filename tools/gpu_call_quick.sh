#!/bin/bash
# quick perf iteration: op-level parity of the GEMM-family kernels, micro-benchmarks, role-wait profile, bench line
set -u
cd "$(dirname "$0")/.."
T=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_r2.py -m gpu -q -x -p no:cacheprovider -k "${KEXPR:-mlp or rowgemm or fused or block or attention}" > $O/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $O/${T}_tests.log
timeout 600 python tools/bench_ops.py --ops ${OPS:-mlp,qkv,proj,fused} > $O/${T}_ops.txt 2>&1
cat $O/${T}_ops.txt
if [ -n "${PROF:-}" ]; then
for cm in "96 1920000" "192 483840" "384 122880"; do
  set -- $cm
  SWN_LIB_VARIANT=prof timeout 120 python tools/mlp_phase_profile.py --C $1 --M $2 >> $O/${T}_mlp_phase.txt 2>&1
done
cat $O/${T}_mlp_phase.txt
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-library-bar --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err
echo "bench rc=$?"; python - <<PY
import json
d = json.load(open("$O/${T}_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity_check", {}).get("ok"), d.get("parity_check", {}).get("max_rel"))
for k, v in d["kernels"].items():
    print(f"  {k:12s} {v['kernel_ms_per_step']:7.2f} ms  share {v['kernel_share_of_step']:.3f}  {v['achieved_tflops']:7.1f} TF/s  {v['achieved_gbs']:7.1f} GB/s  n={v['launches_per_step']}")
print("share sum", d["kernel_share_sum"])
PY
tail -2 $O/${T}_bench.err
