#!/bin/bash
# 2-GPU check (gpurun --gpus 2): TP generate parity on the CUDA kernels, then the bench line at N = 2.
TAG=${1:-r01m}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
echo "=== tp parity" | tee -a $S
timeout -k 10 240 $TR scripts/check_tp_gpu.py > gpurun_out/tp_parity_${TAG}.log 2>&1; echo "exit $? : $(grep -E 'tp=|parity' gpurun_out/tp_parity_${TAG}.log | tr '\n' ' ')" | tee -a $S
echo "=== bench N=2" | tee -a $S
timeout -k 10 300 $TR bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_${TAG}.log 2> gpurun_out/bench_n2_${TAG}.err; echo "exit $? : $(tail -n 1 gpurun_out/bench_n2_${TAG}.log | cut -c1-600)" | tee -a $S
cat $S
