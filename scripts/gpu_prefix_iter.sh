#!/bin/bash
# Fast iteration on the prefix kernel (run under gpurun): parity tests that touch it, bench line, launch list.
TAG=${1:-iter}
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 60 --timeout-method thread"
timeout 150 $PT tests/test_attention_gpu.py -x > gpurun_out/tests_${TAG}.log 2>&1
echo "tests exit $? : $(tail -n 3 gpurun_out/tests_${TAG}.log | tr '\n' ' ')"
timeout 600 python bench.py --e2e-steps 0 --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -c 2500 gpurun_out/bench_${TAG}.json; tail -n 5 gpurun_out/bench_${TAG}.err

