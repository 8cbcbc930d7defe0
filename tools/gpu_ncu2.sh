#!/bin/bash
# ncu --set full capture of selected kernels from a short bench run.  Usage: tools/gpu_ncu2.sh TAG REGEX [skip] [count]
TAG=${1:-x}; RE=${2:-k_deblock2}; SKIP=${3:-6}; CNT=${4:-2}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG python bench.py --streams 128 --max-pictures 10 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
