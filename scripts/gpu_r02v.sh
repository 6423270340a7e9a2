#!/bin/bash
# 8-GPU call: multi-rank tests, all-reduce latency + trace, bench at N = 8, cfg#5 (Llama-2-13B tp=8)
TAG=${1:-r02v}
N=${2:-8}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 14 | cut -c1-1500)" | tee -a $S; }
run tests_multi 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 500 --timeout-method thread tests/test_multigpu_gpu.py -k "$N"
AR_BLOCKS=16,32,64 HYDRAGEN_B200_AR_UNROLL=4 run time_u4 200 $TR scripts/time_allreduce.py
HG_EXTRA_NVCC_FLAGS="-DHG_AR_TRACE" AR_BLOCKS=16,32 run trace 200 $TR scripts/trace_allreduce.py
run bench 600 $TR bench.py --gpus $N --steps 20 --warmup 5
run cfg5 900 $TR scripts/cfg5_tp.py
