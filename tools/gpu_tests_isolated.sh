#!/bin/bash
# Runs every GPU test id in its own process (a trapped kernel poisons the CUDA context, so one failure must not
# mask the others) with a per-test timeout; writes a summary to gpurun_out/isolated_tests.txt.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/isolated_tests.txt
: > $OUT
ids=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
for id in $ids; do
  timeout 300 python -m pytest "$id" -x -q -m gpu > gpurun_out/_one.log 2>&1
  rc=$?
  echo "rc=$rc $id" >> $OUT
  if [ $rc -ne 0 ]; then
    echo "----- $id" >> gpurun_out/isolated_failures.log
    grep -E "Error|error|assert|swn:|relerr|timeout" gpurun_out/_one.log | head -12 >> gpurun_out/isolated_failures.log
  fi
done
echo "passed: $(grep -c '^rc=0' $OUT) failed: $(grep -vc '^rc=0' $OUT)"
grep -v '^rc=0' $OUT | head -60
