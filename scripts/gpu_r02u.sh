#!/bin/bash
# multi-GPU validation (gpurun --gpus N): torchrun pytest (all-reduce parity, TP generate parity), all-reduce latency sweep + stage trace, bench at N
TAG=${1:-r02u}
N=${2:-2}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 14 | cut -c1-1200)" | tee -a $S; }
run tests_multi 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 --timeout-method thread tests/test_multigpu_gpu.py -k "$N"
AR_BLOCKS=16,32,64 HYDRAGEN_B200_AR_UNROLL=4 run time_u4 200 $TR scripts/time_allreduce.py
AR_BLOCKS=16,32,64 HYDRAGEN_B200_AR_UNROLL=2 run time_u2 200 $TR scripts/time_allreduce.py
HG_EXTRA_NVCC_FLAGS="-DHG_AR_TRACE" AR_BLOCKS=16,32 run trace 200 $TR scripts/trace_allreduce.py
run bench 600 $TR bench.py --gpus $N --steps 20 --warmup 5
run bench_ref 400 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5
