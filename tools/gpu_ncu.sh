#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step and one --set full capture of the hot kernels.  Usage: tools/gpu_ncu.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_inter|k_intra|k_bs|k_deblock|k_residual" -s 25 -c 10 -f -o gpurun_out/prof_$TAG python bench.py --streams 128 --max-pictures 10 --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
