#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
grep -E "MemTotal|MemAvailable" /proc/meminfo > gpurun_out/mem_$TAG.txt; nproc >> gpurun_out/mem_$TAG.txt
python bench.py --steps 2 --warmup 3 --no-cpu --no-bitstream > gpurun_out/bench_$TAG.batch.json 2> gpurun_out/bench_$TAG.batch.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-bitstream --host-layout shared > gpurun_out/bench_$TAG.shared.json 2> gpurun_out/bench_$TAG.shared.err
