#!/bin/bash
# multi-GPU (gpurun --gpus N): knob sweep of the fused o_proj + all-reduce launch
TAG=${1:-r02zf}
N=${2:-2}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 50 | cut -c1-300)" | tee -a $S; }
run sweep 200 $TR scripts/sweep_oproj.py
