#!/bin/bash
# e2e (host buffers in, host pictures out) under different read-back paths.  Usage: tools/gpu_pcie_ab.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-bitstream"
$B > gpurun_out/pcie_$TAG.default.json 2> gpurun_out/pcie_$TAG.default.err
for n in 16 64; do
H264B2_D2H_ZEROCOPY=$n $B > gpurun_out/pcie_$TAG.zc$n.json 2> gpurun_out/pcie_$TAG.zc$n.err
done
H264B2_D2H_ZEROCOPY=32 $B --e2e-mode d2h > gpurun_out/pcie_$TAG.zc32_d2honly.json 2> gpurun_out/pcie_$TAG.zc32_d2honly.err
$B --e2e-mode d2h > gpurun_out/pcie_$TAG.default_d2honly.json 2> gpurun_out/pcie_$TAG.default_d2honly.err
