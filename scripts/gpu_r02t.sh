#!/bin/bash
TAG=${1:-r02t}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 12 gpurun_out/${name}_${TAG}.log | cut -c1-1500)" | tee -a $S; }
run tests_all 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread tests
run time_1024 100 python scripts/time_prefix.py
run r01_1024 100 python scripts/time_prefix_r01.py
run hierarchy 200 python scripts/time_hierarchy.py
HYDRAGEN_B200_PREFIX_SPLIT_OVERHEAD=0 run hierarchy_streamk 200 python scripts/time_hierarchy.py
run bench 900 python bench.py --steps 20 --warmup 5
run bench_ref 400 python bench.py --impl reference --steps 20 --warmup 5
