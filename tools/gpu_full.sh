#!/bin/bash
# parity tests + the full default bench line (as the driver runs it) + stream-count variants of the device-timed number.  Usage: tools/gpu_full.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/gpu_tests_$TAG.log
( time timeout 900 python bench.py ) > gpurun_out/full_$TAG.json 2> gpurun_out/full_$TAG.err
for s in 192 256; do timeout 300 python bench.py --streams $s --steps 2 --warmup 2 --no-e2e --no-cpu --no-bitstream --no-configs > gpurun_out/full_$TAG.s$s.json 2> gpurun_out/full_$TAG.s$s.err; done
tail -2 gpurun_out/gpu_tests_$TAG.log
