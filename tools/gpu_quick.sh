#!/bin/bash
# quick job: a few parity tests + device bench variants.  Usage: tools/gpu_quick.sh TAG "pytest -k expr" ["ENV=1" ...]
TAG=${1:-x}; K=${2:-golden}; shift; shift
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -15 > gpurun_out/gpu_tests_$TAG.log
timeout 240 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-bitstream > gpurun_out/ab_$TAG.default.json 2> gpurun_out/ab_$TAG.default.err
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 240 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-bitstream > gpurun_out/ab_$TAG.v$i.json 2> gpurun_out/ab_$TAG.v$i.err
done
tail -3 gpurun_out/gpu_tests_$TAG.log
