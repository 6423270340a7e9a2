#!/bin/bash
# ncu captures of the shipped kernels (full set, with source) + launch list of the bench command
TAG=${1:-r02w}
mkdir -p gpurun_out
prof() { local name=$1; local rx=$2; shift; shift; timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 2 -f -o gpurun_out/prof_${name}_${TAG} "$@" > gpurun_out/ncu_${name}_${TAG}.log 2>&1; echo "$name: ncu exit $?"; }
prof prefix_unit prefix_unit python scripts/ncu_hierarchy.py
prof prefix_grouped prefix_attn_sm100 python scripts/ncu_hierarchy.py
prof decode_slot decode_slot python scripts/ncu_hierarchy.py
ls -la gpurun_out/*${TAG}*
