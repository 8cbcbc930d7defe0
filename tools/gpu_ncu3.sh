#!/bin/bash
# ncu --set full capture of a few launches of selected kernels; the raw and source pages are exported as CSV on the box (reports are
# too large to bring back).  Usage: tools/gpu_ncu3.sh TAG REGEX [skip] [count]
TAG=${1:-x}; RE=${2:-k_intra}; SKIP=${3:-4}; CNT=${4:-2}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o /tmp/prof_$TAG python bench.py --streams 128 --max-pictures 10 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream --no-configs > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_$TAG.raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_$TAG.source.csv 2>/dev/null
ls -la gpurun_out | tail -4
