#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-bitstream"
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/gpu_tests_$TAG.log
$B --host-layout shared > gpurun_out/pm_$TAG.A_ce_shared.json 2> gpurun_out/pm_$TAG.A.err
$B > gpurun_out/pm_$TAG.B_ce_batch.json 2> gpurun_out/pm_$TAG.B.err
H264B2_H2D_ZEROCOPY=64 $B --host-layout shared > gpurun_out/pm_$TAG.D_pull64_shared.json 2> gpurun_out/pm_$TAG.D.err
H264B2_D2H_CHUNK_MB=1024 H264B2_H2D_ZEROCOPY=64 $B --host-layout shared > gpurun_out/pm_$TAG.C_pull64_shared_merged.json 2> gpurun_out/pm_$TAG.C.err
