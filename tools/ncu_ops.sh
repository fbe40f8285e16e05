#!/bin/bash
# ncu --set full of single op launches: tools/ncu_ops.sh tag "ops" only regex skip   (all ncu runs in one gpurun call)
cd "$(dirname "$0")/.."
T=$1; OPS=$2; ONLY=$3; RE=$4; SKIP=${5:-1}
python tools/bench_ops.py --ops $OPS --only $ONLY --once > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 -f -o gpurun_out/$T python tools/bench_ops.py --ops $OPS --only $ONLY --once > gpurun_out/${T}_ncu.log 2>&1
tail -1 gpurun_out/${T}_ncu.log
