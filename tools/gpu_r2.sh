#!/bin/bash
# round-2 A/B job: parity tests, then device-timed bench lines for the listed environment variants.  Usage: tools/gpu_r2.sh TAG ["ENV=1 ENV2=x" ...]
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests_$TAG.log
timeout 240 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-bitstream > gpurun_out/ab_$TAG.default.json 2> gpurun_out/ab_$TAG.default.err
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 240 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-bitstream > gpurun_out/ab_$TAG.v$i.json 2> gpurun_out/ab_$TAG.v$i.err
done
tail -3 gpurun_out/gpu_tests_$TAG.log
