#!/bin/bash
# debugging aid: run one multi-stream decode under compute-sanitizer and keep the report
mkdir -p gpurun_out
cat > /tmp/dbg.py <<'PY'
import sys
sys.path.insert(0, ".")
from h264_video_decoder_demo_b200 import frontend
p = "tests/golden/synth_b_explicit_6x5.first9.h264"
print(frontend.multi_decode([p] * 3, threads=2, readback=True))
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/dbg.py > gpurun_out/sanitizer.log 2>&1
tail -60 gpurun_out/sanitizer.log
