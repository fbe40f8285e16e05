#!/bin/bash
# local helper: rebuild the in-tree libraries (product + bf16), then run a command on the GPU box
cd "$(dirname "$0")/.."
python __graft_entry__.py build 2>&1 | grep -v "offline comp" | grep -v "^$" | tail -3
T=${GRUN_TIMEOUT:-1200}
exec /usr/local/graft/bin/gpurun ${GRUN_GPUS:+--gpus $GRUN_GPUS} --timeout $T -- "$@"
