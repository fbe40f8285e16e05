#!/bin/bash
# Round-2 first GPU call: surrogate training with the reference trainers, full GPU suite, bf16-variant suite,
# bench (parity check + library bar), MLP role-wait profile.  Everything logs into gpurun_out/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O tests/golden/_weights
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
echo "== train" ; date
timeout 900 python tools/train_surrogate.py --out $O/surrogate_wnet_em_fp16.pt > $O/r2a_train.log 2>&1
echo "train rc=$?"; tail -2 $O/r2a_train.log | cut -c1-600
if [ -f $O/surrogate_wnet_em_fp16.pt ]; then
  cp $O/surrogate_wnet_em_fp16.pt tests/golden/_weights/
  cp $O/surrogate_wnet_em_fp16.json tests/golden/surrogate_trained.json
fi
echo "== tests (fp16 product build)"; date
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2a_tests.log 2>&1
echo "tests rc=$?"; tail -15 $O/r2a_tests.log
echo "== bench"; date
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err
echo "bench rc=$?"; cut -c1-1500 $O/r2a_bench.json; tail -3 $O/r2a_bench.err
echo "== bf16 variant suite"; date
SWN_LIB_VARIANT=bf16 timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_r2.py tests/test_gpu_gates.py -m gpu -q -p no:cacheprovider > $O/r2a_tests_bf16.log 2>&1
echo "bf16 rc=$?"; tail -25 $O/r2a_tests_bf16.log
echo "== mlp role-wait profile"; date
for cm in "96 1920000" "192 483840" "384 122880" "384 30720"; do
  set -- $cm
  SWN_LIB_VARIANT=prof timeout 120 python tools/mlp_phase_profile.py --C $1 --M $2 >> $O/r2a_mlp_phase.txt 2>&1
done
cat $O/r2a_mlp_phase.txt
echo "== library bar (tool, B=1/8/64 incl. fp16 autocast)"; date
timeout 600 python tools/library_bar.py --out $O/r2_ref_eager_b200.json > $O/r2a_libbar.log 2>&1
echo "libbar rc=$?"; tail -12 $O/r2a_libbar.log
date
