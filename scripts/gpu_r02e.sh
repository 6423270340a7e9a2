#!/bin/bash
# 1-GPU call: the persistent grouped prefix kernel -- parity tests first, then timing
TAG=${1:-r02e}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 6 gpurun_out/${name}_${TAG}.log | cut -c1-600)" | tee -a $S; }
run tests_attn 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 180 --timeout-method thread -x tests/test_attention_gpu.py
run time_1024 100 python scripts/time_prefix.py
HYDRAGEN_B200_PREFIX_SPLIT=0 run time_1024_nosplit 100 python scripts/time_prefix.py
TP_B=4096 run time_4096 100 python scripts/time_prefix.py
TP_H=4 run time_h4 100 python scripts/time_prefix.py
TP_L=16384 TP_B=2048 TP_H=5 run time_cfg5 100 python scripts/time_prefix.py
run tests_rest 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 180 --timeout-method thread tests --deselect tests/test_attention_gpu.py
run bench 400 python bench.py --steps 50 --warmup 5
