#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/gpu_tests_$TAG.log
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-bitstream"
$B > gpurun_out/la_$TAG.on.json 2> gpurun_out/la_$TAG.on.err
H264B2_LOOKAHEAD=0 $B > gpurun_out/la_$TAG.off.json 2> gpurun_out/la_$TAG.off.err
