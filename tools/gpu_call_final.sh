#!/bin/bash
# round-end evidence: full GPU suite, bench line, ncu pass over one bench step (launch list + DRAM traffic per family)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
if [ -z "${NCU_ONLY:-}" ]; then
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/final_tests.log 2>&1
echo "tests rc=$?"; tail -5 $O/final_tests.log
timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err
echo "bench rc=$?"; tail -2 $O/final_bench.err
fi
# launches of one step (bench.py counts them): the ncu pass skips the 3 warm-up steps and captures exactly one step
L=$(python -c "import json; d = json.load(open('$O/final_bench.json')); print(d['gpu_launches'] // d['steps'])" 2>/dev/null || echo ${LAUNCHES:-220})
echo "launches per step: $L"
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar --no-parity-check > $O/final_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:^(adamw_multi|bilinear_up|bucket_copy|conv_head|copy_cols|cross_attn|dspace_hist|ensure_2ch|expand_warp|minmax_init|mlp|mlp_persist|normalize|patch_embed|rowgemm|rowgemm_persist|sigmoid_mask|swin_attn_stream|swin_block_small|swin_fused|swin_warp_block|window_attn)" -s $((3 * L)) -c $L --csv --log-file $O/final_step.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar --no-parity-check > $O/final_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/final_ncu.log
