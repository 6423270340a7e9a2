#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
S=gpurun_out/summary_${TAG}.txt; : > $S
run() { local name=$1; shift; local t=$1; shift; echo "=== $name" | tee -a $S; timeout -k 10 $t "$@" > gpurun_out/${name}_${TAG}.log 2>&1; echo "exit $? : $(tail -n 34 gpurun_out/${name}_${TAG}.log | cut -c1-220)" | tee -a $S; }
run r01_1024 100 python scripts/time_prefix_r01.py
HYDRAGEN_B200_PREFIX_SPLIT=0 run time_1024_nosplit 100 python scripts/time_prefix.py
HYDRAGEN_B200_PREFIX_DBG=8 HYDRAGEN_B200_PREFIX_SPLIT=0 run time_nosplit_dbg8 100 python scripts/time_prefix.py
HYDRAGEN_B200_PREFIX_CTAS=64 HYDRAGEN_B200_PREFIX_SPLIT=0 run time_nosplit_64ctas 100 python scripts/time_prefix.py
HYDRAGEN_B200_PDL=0 HYDRAGEN_B200_PREFIX_SPLIT=0 run time_nosplit_nopdl 100 python scripts/time_prefix.py
HYDRAGEN_B200_PDL=0 run r01_nopdl 100 python scripts/time_prefix_r01.py
HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" HYDRAGEN_B200_PREFIX_SPLIT=0 run trace_nosplit 100 python scripts/trace_prefix.py
HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" HYDRAGEN_B200_PREFIX_SPLIT=0 HYDRAGEN_B200_PREFIX_CTAS=64 run trace_nosplit_64 100 python scripts/trace_prefix.py
HG_EXTRA_NVCC_FLAGS="-DHG_PREFIX_TRACE" HYDRAGEN_B200_PREFIX_SPLIT=0 HYDRAGEN_B200_PDL=0 run trace_nosplit_nopdl 100 python scripts/trace_prefix.py
