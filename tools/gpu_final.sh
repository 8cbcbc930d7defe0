#!/bin/bash
# round-end evidence: parity tests, smoke(), the full bench line (as the driver runs it), the reference arm, and the ncu launch list
# of one step of the same command.  Usage: tools/gpu_final.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/final_$TAG.tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/final_$TAG.tests.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/final_$TAG.bench.json 2> gpurun_out/final_$TAG.bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_$TAG.reference.json 2> gpurun_out/final_$TAG.reference.err
# launch list: skip the parity-checked warm-up pass (76 submits x ~8 kernels), list one timed step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/final_$TAG.launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-bitstream --no-configs > gpurun_out/final_$TAG.ncu_launch.log 2>&1
tail -3 gpurun_out/final_$TAG.tests.log
