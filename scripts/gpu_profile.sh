#!/bin/bash
# bench + ncu evidence (run under gpurun, 1 GPU): bench line, launch list, one --set full capture per hot kernel.
# usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
BENCH_ARGS="--steps 2 --warmup 3 --no-graph --e2e-steps 0 --no-cpu-baseline"
echo "=== bench" | tee gpurun_out/profile_summary.txt
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "exit $?" | tee -a gpurun_out/profile_summary.txt
tail -c 3000 gpurun_out/bench_${TAG}.json | tee -a gpurun_out/profile_summary.txt
echo "=== launches" | tee -a gpurun_out/profile_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prefix_attn|rowwise_attn|kv_append|combine|decode_slot" -s 200 -c 400 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py $BENCH_ARGS > gpurun_out/ncu_launches.log 2>&1
echo "exit $?" | tee -a gpurun_out/profile_summary.txt
for K in prefix_attn decode_slot; do
  echo "=== full $K" | tee -a gpurun_out/profile_summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 100 -c 2 -f -o gpurun_out/prof_${K}_${TAG} \
    python bench.py $BENCH_ARGS > gpurun_out/ncu_${K}.log 2>&1
  echo "exit $?" | tee -a gpurun_out/profile_summary.txt
done
ls -la gpurun_out | tee -a gpurun_out/profile_summary.txt
