"""Parser-thread sweep of the whole-decoder pipeline (h264b2_multi_decode) on one GPU: frames/s against the number of parser threads.
usage: python tools/threads_sweep.py [streams]   (on the GPU box; prints one JSON line per thread count)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from h264_video_decoder_demo_b200 import frontend

nstr = int(sys.argv[1]) if len(sys.argv) > 1 else 64
f = os.path.join(bench.REF_DIR, "streams", bench.WORKLOADS["B_frames.cabac"] + ".h264")
paths = [f] * nstr
frontend.multi_decode(paths[:4], device=0, threads=4, readback=True, hashes=False)
cores = bench.usable_cores()
hashes = os.environ.get("SWEEP_HASHES", "0") == "1"
ts = [int(x) for x in os.environ["SWEEP_THREADS"].split(",")] if os.environ.get("SWEEP_THREADS") else (cores // 2, cores - 4, cores - 2, cores - 1, cores, cores + 2, cores + 4, 2 * cores)
for t in ts:
    if t < 1:
        continue
    st, _ = frontend.multi_decode(paths, device=0, threads=t, readback=True, hashes=hashes)
    print(json.dumps({"threads": t, "queue_depth": os.environ.get("H264B2_MULTI_QUEUE_DEPTH", "3"), "hashes": hashes, "cores": cores, "streams": nstr, "frames_per_s": round(st["frames_out"] / st["seconds"], 1),
                      "host_stage_pictures_per_s_per_thread": round(st["pictures"] / st["parse_seconds"], 1) if st["parse_seconds"] > 0 else None}), flush=True)
