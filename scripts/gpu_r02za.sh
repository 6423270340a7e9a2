#!/bin/bash
# multi-GPU (gpurun --gpus N): stage trace of the fused o_proj + all-reduce launch
TAG=${1:-r02za}
N=${2:-2}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 30 | cut -c1-600)" | tee -a $S; }
HG_EXTRA_NVCC_FLAGS="-DHG_OPROJ_TRACE" HYDRAGEN_B200_OPROJ_BN=128 run trace_bn128 150 $TR scripts/trace_oproj.py
HG_EXTRA_NVCC_FLAGS="-DHG_OPROJ_TRACE" HYDRAGEN_B200_OPROJ_BN=256 run trace_bn256 150 $TR scripts/trace_oproj.py
