#!/bin/bash
# One box visit for the round-end evidence of the current sources: the ncu --set full traffic capture (tools/gpu_traffic.sh), the traffic
# file rebuilt from it on the box so that the bench line that follows carries roofline.traffic, then tools/gpu_final.sh.
# Usage: tools/gpu_evidence.sh TAG      (afterwards, here: python tools/make_traffic.py gpurun_out/prof_TAG.raw.csv gpurun_out/prof_TAG.libsha1)
TAG=${1:-r02}
tools/gpu_traffic.sh $TAG
python tools/make_traffic.py gpurun_out/prof_$TAG.raw.csv gpurun_out/prof_$TAG.libsha1 > gpurun_out/make_traffic_$TAG.log 2>&1
tools/gpu_final.sh $TAG
