#!/bin/bash
TAG=${1:-r02n}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 14 gpurun_out/${name}_${TAG}.log | cut -c1-220)" | tee -a $S; }
run tests_attn 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 180 --timeout-method thread -x tests/test_attention_gpu.py
run r01_1024 100 python scripts/time_prefix_r01.py
run time_1024 100 python scripts/time_prefix.py
HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" run trace_new 100 python scripts/trace_prefix.py
