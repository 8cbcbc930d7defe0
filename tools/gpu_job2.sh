#!/bin/bash
# parity tests, the full bench line (packed coefficient transport), the same e2e with dense coefficients.  Usage: tools/gpu_job2.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests_$TAG.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-bitstream --dense-coefs > gpurun_out/bench_$TAG.dense.json 2> gpurun_out/bench_$TAG.dense.err
