#!/bin/bash
# A/B of an alternative engine build: parity tests and the device bench with H264B2_LIB.  Usage: tools/gpu_job8.sh TAG lib
TAG=${1:-x}; LIB=$PWD/$2
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-bitstream --no-e2e"
$B > gpurun_out/alt_$TAG.default.json 2> gpurun_out/alt_$TAG.default.err
H264B2_LIB=$LIB timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/alt_$TAG.tests.log
H264B2_LIB=$LIB timeout 300 $B > gpurun_out/alt_$TAG.alt.json 2> gpurun_out/alt_$TAG.alt.err
