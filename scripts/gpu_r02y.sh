#!/bin/bash
# 1 GPU: the o_proj tcgen05 GEMM alone (world == 1) -- parity tests and timing against cuBLAS
TAG=${1:-r02y}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 30 | cut -c1-600)" | tee -a $S; }
run tests_oproj 300 python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread tests/test_oproj_gpu.py
run time_oproj 200 python scripts/time_oproj.py
