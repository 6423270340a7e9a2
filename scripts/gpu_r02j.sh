#!/bin/bash
# ncu A/B: round-1 kernel vs persistent kernel (whole units / stream-K), 2 launches each, full set with source
TAG=${1:-r02j}
mkdir -p gpurun_out
prof() { local name=$1; shift; timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefix_attn -s 24 -c 2 -f -o gpurun_out/prof_${name}_${TAG} "$@" > gpurun_out/ncu_${name}_${TAG}.log 2>&1; echo "$name: ncu exit $?"; tail -n 2 gpurun_out/ncu_${name}_${TAG}.log; }
prof r01 python scripts/time_prefix_r01.py
HYDRAGEN_B200_PREFIX_SPLIT=0 prof nosplit python scripts/time_prefix.py
prof split python scripts/time_prefix.py
ls -la gpurun_out/*.ncu-rep
