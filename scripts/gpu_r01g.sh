#!/bin/bash
# r01g GPU call: all -m gpu tests; sweep of the tile-B softmax phase offset (HYDRAGEN_B200_PREFIX_BDELAY, cycles);
# clock64 trace with an offset.
TAG=${1:-r01g}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 3 gpurun_out/${name}_${TAG}.log | tr '\n' ' ' | cut -c1-700)" | tee -a $S; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
run tests 200 $PT tests/test_attention_gpu.py
for D in 0 300 600 900 1200 1500 2000; do
  HYDRAGEN_B200_PREFIX_BDELAY=$D run time_d${D} 100 python scripts/time_prefix.py
done
for D in 600 1200; do
  HYDRAGEN_B200_PREFIX_BDELAY=$D TP_B=4096 run time4096_d${D} 100 python scripts/time_prefix.py
done
HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_TRACE HYDRAGEN_B200_PREFIX_BDELAY=900 run trace_d900 120 python scripts/trace_prefix.py
cat $S
