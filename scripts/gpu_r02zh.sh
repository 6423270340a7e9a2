#!/bin/bash
# 1 GPU, end of round 2: the whole GPU suite, smoke(), the bench line, ncu capture of the o_proj GEMM
TAG=${1:-r02zh}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 12 gpurun_out/${name}_${TAG}.log | cut -c1-1500)" | tee -a $S; }
run tests_all 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread tests
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench 900 python bench.py --steps 20 --warmup 5
run ncu_oproj 300 ncu --set full --clock-control none --import-source on -k regex:oproj -s 2 -c 4 -o gpurun_out/oproj_gemm_${TAG} -f python scripts/ncu_oproj.py
