#!/bin/bash
TAG=${1:-r02o}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 32 gpurun_out/${name}_${TAG}.log | cut -c1-220)" | tee -a $S; }
HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" run trace_new 100 python scripts/trace_prefix.py
