#!/bin/bash
# one ncu --set full capture of a single op launch:  tools/gpu_call_ncu2.sh tag ops only kernel-regex skip
cd "$(dirname "$0")/.."
bash tools/ncu_ops.sh "$@"
