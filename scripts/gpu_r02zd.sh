#!/bin/bash
# multi-GPU (gpurun --gpus N): fused o_proj + all-reduce -- tests (TP generate through it, parity + stress), timing, trace, bench with the oproj_allreduce key
TAG=${1:-r02zd}
N=${2:-2}
MODE=${3:-full}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 16 | cut -c1-1500)" | tee -a $S; }
run tests_multi 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 400 --timeout-method thread tests/test_multigpu_gpu.py -k "$N and (tp_generate or fused)"
if [ "$MODE" = full ]; then
  OPROJ_SHAPES="1024,4096,4096;2048,5120,5120" run check_oproj 150 $TR scripts/check_oproj_allreduce.py
  HYDRAGEN_B200_OPROJ_RWARPS=4 OPROJ_FUSED_ONLY=1 OPROJ_SHAPES="1024,4096,4096;2048,5120,5120" run check_rw4 150 $TR scripts/check_oproj_allreduce.py
  HG_EXTRA_NVCC_FLAGS="-DHG_OPROJ_TRACE" run trace 150 $TR scripts/trace_oproj.py
fi
run bench 600 $TR bench.py --gpus $N --steps ${BENCH_STEPS:-20} --warmup 5
