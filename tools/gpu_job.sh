#!/bin/bash
# One GPU-box job: parity tests, the bench line, parser-thread scaling of the Annex-B pipeline.  Usage: tools/gpu_job.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests_$TAG.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err >> gpurun_out/gpu_tests_$TAG.log
for t in 1 4 8 16; do
python - $t >> gpurun_out/multi_scaling_$TAG.jsonl 2>> gpurun_out/multi_scaling_$TAG.err <<'PY'
import sys, json, os
sys.path.insert(0, ".")
from h264_video_decoder_demo_b200 import frontend
t = int(sys.argv[1])
p = [os.path.join("oracle/_ref/streams", "HeavyHand_1080p.B_frames.cabac.h264")] * 32
st, _ = frontend.multi_decode(p, threads=t, readback=True, hashes=False)
print(json.dumps({"fps": round(st["frames_out"] / st["seconds"], 1), **st}))
PY
done
