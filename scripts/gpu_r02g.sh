#!/bin/bash
TAG=${1:-r02g}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 14 gpurun_out/${name}_${TAG}.log | cut -c1-300)" | tee -a $S; }
run tests_attn 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 180 --timeout-method thread -x tests/test_attention_gpu.py
run r01_1024 100 python scripts/time_prefix_r01.py
run time_1024 100 python scripts/time_prefix.py
HYDRAGEN_B200_PREFIX_SPLIT=0 run time_1024_nosplit 100 python scripts/time_prefix.py
run r01_1024_again 100 python scripts/time_prefix_r01.py
TP_B=4096 run r01_4096 100 python scripts/time_prefix_r01.py
TP_B=4096 run time_4096 100 python scripts/time_prefix.py
TP_H=4 run time_h4 100 python scripts/time_prefix.py
TP_L=16384 TP_B=2048 TP_H=5 run r01_cfg5 100 python scripts/time_prefix_r01.py
TP_L=16384 TP_B=2048 TP_H=5 run time_cfg5 100 python scripts/time_prefix.py
HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" run trace_split 100 python scripts/trace_prefix.py
HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" HYDRAGEN_B200_PREFIX_SPLIT=0 run trace_nosplit 100 python scripts/trace_prefix.py
HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" TP_H=4 run trace_h4 100 python scripts/trace_prefix.py
