#!/bin/bash
# Full GPU check (run under gpurun): every -m gpu test, smoke(), kernel-level quick bench; logs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name timeout cmd...
  local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit $? : $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400)" | tee -a gpurun_out/summary.txt
}
run tests 400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 90 --timeout-method thread
run smoke 300 python __graft_entry__.py smoke
run microbench 300 python scripts/microbenchmark.py --pairs 1024:2048 --suffix 1,64
cat gpurun_out/summary.txt
