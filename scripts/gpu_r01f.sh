#!/bin/bash
# r01f GPU call: all -m gpu tests (base softmax), graph-timed prefix kernel base vs split-column softmax,
# parity tests again under the split form, short bench under each.
TAG=${1:-r01f}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 3 gpurun_out/${name}_${TAG}.log | tr '\n' ' ' | cut -c1-700)" | tee -a $S; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 90 --timeout-method thread"
run tests_base 420 $PT tests
run time_base_1024 100 python scripts/time_prefix.py
TP_B=4096 run time_base_4096 100 python scripts/time_prefix.py
export HYDRAGEN_B200_PREFIX_SOFTMAX=split
run time_split_1024 100 python scripts/time_prefix.py
TP_B=4096 run time_split_4096 100 python scripts/time_prefix.py
TP_B=128 run time_split_128 100 python scripts/time_prefix.py
run tests_split 300 $PT tests/test_attention_gpu.py tests/test_llama_gpu.py
run bench_split 300 python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline
unset HYDRAGEN_B200_PREFIX_SOFTMAX
run bench_base 300 python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline
cat $S
