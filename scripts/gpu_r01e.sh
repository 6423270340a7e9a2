#!/bin/bash
# r01e GPU call: all -m gpu tests, smoke, MUFU issue microbenchmark, prefix-kernel clock64 traces (two batch sizes),
# graph-timed prefix kernel at three batch sizes, one ncu --set full capture (with source) of the prefix kernel, short bench.
TAG=${1:-r01e}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 3 gpurun_out/${name}_${TAG}.log | tr '\n' ' ' | cut -c1-600)" | tee -a $S; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
run tests 420 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 90 --timeout-method thread
run smoke 200 python __graft_entry__.py smoke
run mufu 60 scripts/microbench/mufu_issue
HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_TRACE run trace1024 120 python scripts/trace_prefix.py
HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_TRACE TRACE_B=128 run trace128 120 python scripts/trace_prefix.py
run time1024 120 python scripts/time_prefix.py
TP_B=128 run time128 120 python scripts/time_prefix.py
TP_B=4096 run time4096 120 python scripts/time_prefix.py
echo "=== ncu prefix" | tee -a $S
timeout 400 ncu --set full --clock-control none --import-source on -k regex:prefix_attn -s 20 -c 1 -f -o gpurun_out/prof_prefix_${TAG} \
  python scripts/time_prefix.py > gpurun_out/ncu_prefix_${TAG}.log 2>&1
echo "exit $?" | tee -a $S
run bench 400 python bench.py --steps 100 --warmup 10
cat $S
