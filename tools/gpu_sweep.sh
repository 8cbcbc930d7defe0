#!/bin/bash
# stream-count sweep of the device-timed bench.  Usage: tools/gpu_sweep.sh TAG "4 32 128" [ENV=..]
TAG=${1:-x}; LIST=${2:-"4 32 128"}; shift; shift
mkdir -p gpurun_out
for s in $LIST; do
  env "$@" timeout 240 python bench.py --streams $s --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream > gpurun_out/sweep_$TAG.$s.json 2> gpurun_out/sweep_$TAG.$s.err
done
