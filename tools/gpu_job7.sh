#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
for S in 96 160 192; do
python bench.py --steps 2 --warmup 3 --no-cpu --no-bitstream --no-e2e --streams $S > gpurun_out/s_$TAG.$S.json 2> gpurun_out/s_$TAG.$S.err
done
