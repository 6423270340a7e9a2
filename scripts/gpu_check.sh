#!/bin/bash
# Staged GPU check (run under gpurun): each stage has its own timeout and log under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
run() { # name timeout cmd...
  local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit $? : $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
run combine 300 $PT tests/test_combine_lse_gpu.py -x
run rowwise 600 $PT tests/test_attention_gpu.py -x -k "seqlen or causal or kv_append or (golden and rowwise)"
run tcgen05 600 $PT tests/test_attention_gpu.py -k "tcgen05 or rescale"
run rest 900 $PT tests/test_attention_gpu.py -k "not (seqlen or causal or kv_append or (golden and rowwise) or tcgen05 or rescale)"
run smoke 300 python __graft_entry__.py smoke
run quick_bench 300 python scripts/quick_bench.py
"$@"
cat gpurun_out/summary.txt
