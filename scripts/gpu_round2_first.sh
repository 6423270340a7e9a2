#!/bin/bash
# First GPU call of round 2 (run under gpurun, 1 GPU, ~2 min): the alternate-block softmax variant of the prefix kernel
# (HYDRAGEN_B200_PREFIX_SOFTMAX=alt) was written after round 1's GPU budget was spent and has NOT run on hardware.
#   1. parity of the opt-in variants inside pytest (split is always on; simple and alt need the env switch)
#   2. graph-timed prefix kernel: base vs alt at B = 1024 and 4096 (cfg#2: base 32.0-33.0 us; isolated-stream estimate for alt: ~26 us)
#   3. the softmax-stream and MUFU microbenchmarks again (reference points: profiles/r01r_*, profiles/r01e_mufu_issue.txt)
# If alt passes 1. and wins 2.: set HG_PREFIX_SOFTMAX_DEFAULT to 3 in csrc/prefix_sm100.cu, drop the skipif of the 'alt' test
# parameter in tests/test_attention_gpu.py, re-run the full -m gpu suite and bench.py.
TAG=${1:-r02a}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 3 gpurun_out/${name}_${TAG}.log | tr '\n' ' ' | cut -c1-700)" | tee -a $S; }
HYDRAGEN_B200_TEST_EXPERIMENTAL=1 run tests_variants 300 python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread tests/test_attention_gpu.py -k softmax_variant
run time_base_1024 100 python scripts/time_prefix.py
HYDRAGEN_B200_PREFIX_SOFTMAX=alt run time_alt_1024 100 python scripts/time_prefix.py
TP_B=4096 run time_base_4096 100 python scripts/time_prefix.py
HYDRAGEN_B200_PREFIX_SOFTMAX=alt TP_B=4096 run time_alt_4096 100 python scripts/time_prefix.py
HYDRAGEN_B200_PREFIX_SOFTMAX=alt TP_B=128 run time_alt_128 100 python scripts/time_prefix.py
[ -x scripts/microbench/softmax_stream ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/microbench/softmax_stream scripts/microbench/softmax_stream.cu
run softmax_stream 60 scripts/microbench/softmax_stream
cat $S
