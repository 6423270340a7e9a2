#!/bin/bash
# multi-GPU (gpurun --gpus N), end of round 2: multi-rank tests and/or the bench line at N
TAG=${1:-r02zj}
N=${2:-2}
WHAT=${3:-both}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(grep -vE 'Warning|warn|^$|OMP_NUM|\*\*\*' gpurun_out/${name}_${TAG}.log | tail -n 12 | cut -c1-1500)" | tee -a $S; }
if [ "$WHAT" != bench ]; then run tests_multi 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 400 --timeout-method thread tests/test_multigpu_gpu.py -k "$N"; fi
if [ "$WHAT" != tests ]; then run bench 600 $TR bench.py --gpus $N --steps 20 --warmup 5; fi
