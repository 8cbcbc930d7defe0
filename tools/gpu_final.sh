#!/bin/bash
# round-end evidence: parity tests, smoke(), the full bench line, the reference arm.  Usage: tools/gpu_final.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/final_$TAG.tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/final_$TAG.tests.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/final_$TAG.bench.json 2> gpurun_out/final_$TAG.bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_$TAG.reference.json 2> gpurun_out/final_$TAG.reference.err
