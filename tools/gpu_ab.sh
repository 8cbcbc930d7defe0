#!/bin/bash
# A/B job: parity tests, then the device-timed bench line with the default engine and with alternative builds under build/.  Usage: tools/gpu_ab.sh TAG [lib ...]
TAG=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/gpu_tests_$TAG.log
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-bitstream > gpurun_out/ab_$TAG.default.json 2> gpurun_out/ab_$TAG.default.err
for lib in "$@"; do
  n=$(basename $lib .so)
  H264B2_LIB=$PWD/$lib python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-bitstream > gpurun_out/ab_$TAG.$n.json 2> gpurun_out/ab_$TAG.$n.err
done
