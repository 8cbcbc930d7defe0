#!/bin/bash
# device-timed bench under several environment variants.  Usage: tools/gpu_env_ab.sh TAG "ENV=1 ENV2=x" ["..."]
TAG=${1:-x}; shift
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 240 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-bitstream > gpurun_out/env_$TAG.v$i.json 2> gpurun_out/env_$TAG.v$i.err
done
