#!/bin/bash
# r01l GPU call: rope kernel with packed intermediate rounding -- parity tests (bit-exact) and the bench line.
TAG=${1:-r01l}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2> gpurun_out/${name}_${TAG}.err; echo "exit $? : $(tail -n 3 gpurun_out/${name}_${TAG}.log | tr '\n' ' ' | cut -c1-900)" | tee -a $S; }
run tests 300 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread
run bench 400 python bench.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rope_qk -s 40 -c 2 -f -o gpurun_out/prof_rope_qk_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-graph --e2e-steps 0 --no-cpu-baseline > gpurun_out/ncu_rope_qk_${TAG}.log 2>&1
cat $S
