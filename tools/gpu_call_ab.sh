#!/bin/bash
# A/B micro-benchmarks of library variants in one GPU call:  VARIANTS="'' x1 x2" OPS=mlp,fused ONLY=0 bash tools/gpu_call_ab.sh tag
set -u
cd "$(dirname "$0")/.."
T=${1:-ab}
O=gpurun_out
mkdir -p $O
for v in ${VARIANTS:-"_ x1"}; do
  vv=$v; [ "$v" = "_" ] && vv=""
  echo "=== variant '$vv'" | tee -a $O/${T}_ops.txt
  if [ -n "${TESTS:-}" ]; then
    SWN_LIB_VARIANT=$vv timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_r2.py -m gpu -q -x -p no:cacheprovider -k "$TESTS" 2>&1 | tail -3
  fi
  SWN_LIB_VARIANT=$vv timeout 300 python tools/bench_ops.py --ops ${OPS:-mlp} ${ONLY:+--only $ONLY} 2>&1 | tee -a $O/${T}_ops.txt
done
