#!/bin/bash
# r01i GPU call: register split 56/224 (no spills in the softmax loop) x tile-B phase offset sweep; attention parity tests.
TAG=${1:-r01i}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 3 gpurun_out/${name}_${TAG}.log | tr '\n' ' ' | cut -c1-700)" | tee -a $S; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
run tests 200 $PT tests/test_attention_gpu.py
for D in 0 400 600 800 1000 1300; do
  HYDRAGEN_B200_PREFIX_BDELAY=$D run time_d${D} 100 python scripts/time_prefix.py
done
for D in 0 800; do
  HYDRAGEN_B200_PREFIX_BDELAY=$D TP_B=4096 run time4096_d${D} 100 python scripts/time_prefix.py
done
cat $S
