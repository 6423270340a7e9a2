#!/bin/bash
# multi-GPU call of round 2 (gpurun --gpus N): the epoch-barrier all-reduce -- parity vs NCCL, then latency swept over
# CTA count and reductions in flight per thread; stage trace (instrumented build); bench line at N.
TAG=${1:-r02b}
N=${2:-2}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 12 | cut -c1-400)" | tee -a $S; }
run check 200 $TR scripts/check_allreduce.py
AR_BLOCKS=16,32,64 HYDRAGEN_B200_AR_UNROLL=4 run time_u4 200 $TR scripts/time_allreduce.py
AR_BLOCKS=16,32,64 HYDRAGEN_B200_AR_UNROLL=2 run time_u2 200 $TR scripts/time_allreduce.py
AR_BLOCKS=16,32,64 HYDRAGEN_B200_AR_UNROLL=8 run time_u8 200 $TR scripts/time_allreduce.py
HG_EXTRA_NVCC_FLAGS="-DHG_AR_TRACE" AR_BLOCKS=16,32 run trace 200 $TR scripts/trace_allreduce.py
[ -n "$SKIP_BENCH" ] || run bench 300 $TR bench.py --gpus $N --steps 50 --warmup 5
