#!/bin/bash
# Round-end evidence (run under gpurun, 1 GPU): every -m gpu test, smoke(), the default bench line (+ whole-model decode),
# the reference arm, the ncu launch list of the bench command and one --set full capture per hot kernel.
# usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-r01k}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2> gpurun_out/${name}_${TAG}.err; echo "exit $? : $(tail -n 3 gpurun_out/${name}_${TAG}.log | tr '\n' ' ' | cut -c1-900)" | tee -a $S; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
run tests 420 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread
run smoke 200 python __graft_entry__.py smoke
run bench 600 python bench.py --full-model
run bench_reference 300 python bench.py --impl reference --steps 5 --warmup 2
BENCH_ARGS="--steps 2 --warmup 3 --no-graph --e2e-steps 0 --no-cpu-baseline"
echo "=== launches" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prefix_attn|rowwise_attn|kv_append|combine|decode_slot|rope_qk" -s 200 -c 400 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py $BENCH_ARGS > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "exit $?" | tee -a $S
for K in prefix_attn decode_slot rope_qk; do
  echo "=== full $K" | tee -a $S
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 2 -f -o gpurun_out/prof_${K}_${TAG} \
    python bench.py $BENCH_ARGS > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  echo "exit $?" | tee -a $S
done
ls -la gpurun_out | tail -n 30 >> $S
cat $S
