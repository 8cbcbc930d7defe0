#!/bin/bash
# device-timed bench with alternative engine builds.  Usage: tools/gpu_libs.sh TAG STREAMS lib.so [lib.so ...]   ("default" = the in-tree build)
TAG=${1:-x}; S=${2:-128}; shift; shift
mkdir -p gpurun_out
for lib in "$@"; do
  n=$(basename $lib .so)
  if [ "$lib" = default ]; then
    timeout 240 python bench.py --streams $S --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream > gpurun_out/lib_$TAG.$n.json 2> gpurun_out/lib_$TAG.$n.err
  else
    H264B2_LIB=$PWD/$lib timeout 240 python bench.py --streams $S --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream > gpurun_out/lib_$TAG.$n.json 2> gpurun_out/lib_$TAG.$n.err
  fi
done
