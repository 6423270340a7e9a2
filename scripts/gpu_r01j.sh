#!/bin/bash
# r01j GPU call: polynomial-exp2 variants (HG_PREFIX_EMU_EVERY = 4, 2) of the current kernel, with and without the tile-B offset.
TAG=${1:-r01j}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 3 gpurun_out/${name}_${TAG}.log | tr '\n' ' ' | cut -c1-700)" | tee -a $S; }
for E in 4 2; do
  for D in 0 800; do
    HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_EMU_EVERY=$E HYDRAGEN_B200_PREFIX_BDELAY=$D run time_emu${E}_d${D} 100 python scripts/time_prefix.py
  done
done
HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_EMU_EVERY=4 TP_B=4096 run time4096_emu4 100 python scripts/time_prefix.py
HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_EMU_EVERY=4 run tests_emu4 200 python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 tests/test_attention_gpu.py -k "golden or operator or prefix"
cat $S
