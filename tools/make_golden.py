"""Slice short prefixes of the reference-generated replay containers (oracle/_ref/replay/*.bin.xz, made by
oracle/ref_harness from the UNMODIFIED reference) into small committed fixtures under tests/golden/.
Each fixture carries the reference decoder's own pre-/post-deblock picture checksums, so it pins both the
CPU oracle and the CUDA path without /root/reference.   Usage: python tools/make_golden.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h264_video_decoder_demo_b200 import replay  # noqa: E402

# (stream, pictures): tff needs 7 so that the one picture with field macroblocks (decode index 6) is covered
PLAN = [("HeavyHand_1080p.B_frames.cabac", 4), ("HeavyHand_1080p.B_frames_4.no_cabac.no_tff", 3),
        ("HeavyHand_1080p.B_frames_cabac_tff", 7), ("HeavyHand_1080p.no_B_frames.cabac.no_tff", 3), ("gop121.naluCnt453", 3)]

if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, n in PLAN:
        rp = replay.load_replay(os.path.join(replay.default_replay_dir(), name + ".bin.xz"), n)
        dst = os.path.join(out_dir, f"{name}.first{n}.rp.xz")
        replay.save_replay(rp, dst, n, preset=9)
        print(dst, os.path.getsize(dst))
