#!/bin/bash
# one ncu --set full capture (with source) of the prefix kernel at cfg#2; usage: gpu_ncu_prefix.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prefix_attn -s 40 -c 1 -f -o gpurun_out/prof_prefix_${TAG} \
  python bench.py --steps 2 --warmup 3 --no-graph --e2e-steps 0 --no-cpu-baseline > gpurun_out/ncu_prefix_${TAG}.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_prefix_${TAG}.log
