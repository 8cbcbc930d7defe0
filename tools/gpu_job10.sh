#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_api.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/dec_$TAG.tests.log
