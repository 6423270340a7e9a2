#!/bin/bash
# 2-GPU call: per-stage timestamps of the all-reduce kernel (instrumented dev build)
TAG=${1:-r02c}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
HG_EXTRA_NVCC_FLAGS="-DHG_AR_TRACE" AR_BLOCKS=16,64,128 timeout -k 10 200 $TR scripts/trace_allreduce.py > gpurun_out/ar_trace_${TAG}.log 2>&1
grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/ar_trace_${TAG}.log | tail -40
