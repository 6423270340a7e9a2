#!/bin/bash
# multi-GPU (gpurun --gpus N): fused o_proj + all-reduce, stress of the tile-flag ordering variants
TAG=${1:-r02zc}
N=${2:-2}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | grep -E 'stress|world|parity|FAILED' | cut -c1-300)" | tee -a $S; }
for sig in 0 1 2; do for rw in 2 4; do
  HYDRAGEN_B200_OPROJ_SIGNAL=$sig HYDRAGEN_B200_OPROJ_RWARPS=$rw OPROJ_STRESS=40 OPROJ_FUSED_ONLY=1 OPROJ_SHAPES="1024,4096,4096;2048,5120,5120" run stress_sig${sig}_rw$rw 150 $TR scripts/check_oproj_allreduce.py
done; done
