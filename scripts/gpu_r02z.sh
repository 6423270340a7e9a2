#!/bin/bash
# multi-GPU (gpurun --gpus N): o_proj GEMM fused with its all-reduce -- parity against matmul + NCCL, timing against cuBLAS + NVLS kernel / NCCL
TAG=${1:-r02z}
N=${2:-2}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 30 | cut -c1-600)" | tee -a $S; }
run check_oproj 150 $TR scripts/check_oproj_allreduce.py
HYDRAGEN_B200_OPROJ_BN=256 OPROJ_SHAPES="1024,4096,4096;2048,5120,5120" run check_bn256 150 $TR scripts/check_oproj_allreduce.py
HYDRAGEN_B200_OPROJ_BN=128 OPROJ_SHAPES="1024,4096,4096;2048,5120,5120" run check_bn128 150 $TR scripts/check_oproj_allreduce.py
