"""The multi-stream host pipeline (parser thread pool with dynamic stream claiming, per-stream queues, page-locked block pool, submit
thread: csrc/host/h264_multi.cpp + h264_front.cpp) under ThreadSanitizer, on the CPU, against a test double of the engine's C ABI
(tests/mock_engine.cpp — no pixels, it only reads the submitted arrays the way the copy engine would, H264B2_SUBMIT_DEPTH deep).
What it pins: no data race between parser threads, the submit thread's block releases (round 1's advisor finding) and the recycling of
picture blocks; every stream's pictures all arrive, in the same order for every replica and every thread count."""
import glob
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(ROOT, "h264_video_decoder_demo_b200", "csrc", "host")


@pytest.fixture(scope="module")
def tsan_exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("tsan")
    exe = str(d / "multi_pool_tsan")
    srcs = [os.path.join(HERE, "multi_pool_main.cpp"), os.path.join(HERE, "mock_engine.cpp")] + [os.path.join(HOST, f) for f in ("h264_multi.cpp", "h264_front.cpp", "h264_slice.cpp", "h264_params.cpp")]
    r = subprocess.run(["g++", "-O1", "-g", "-fsanitize=thread", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", HOST, "-o", exe] + srcs + ["-pthread"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build not available: " + r.stderr[-300:])
    return exe


def _run(exe, threads, replicas, flags, files, env=None):
    r = subprocess.run([exe, str(threads), str(replicas), str(flags)] + files, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0, r.stderr[-1000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_parser_pool_is_race_free_and_order_preserving(tsan_exe):
    files = sorted(glob.glob(os.path.join(HERE, "golden", "synth_*.h264")))
    if not files:
        pytest.skip("no Annex-B fixtures")
    by_name = {os.path.basename(f): f for f in files}
    picked = [by_name[n] for n in ("synth_b_direct.first7.h264", "synth_mmco_lt_36.first9.h264") if n in by_name] or files[:2]
    for f in picked:                                     # all streams of one context share the picture size: replicas of one file per run
        base = _run(tsan_exe, 1, 5, 1, [f])
        assert base["pictures"] == base["frames_out"] > 0 and len(set(base["hashes"])) == 1
        for threads, env in ((2, None), (4, None), (5, {"H264B2_MULTI_QUEUE_DEPTH": "1"}), (3, {"H264B2_MULTI_QUEUE_DEPTH": "8"})):
            got = _run(tsan_exe, threads, 5, 1, [f], env)
            assert got["hashes"] == base["hashes"] and got["pictures"] == base["pictures"] and got["frames_out"] == base["frames_out"]


def test_closed_gop_units_under_tsan(tsan_exe):
    f = os.path.join(HERE, "golden", "HeavyHand_1080p.no_B_frames.cabac.no_tff.first3.h264")
    if not os.path.exists(f):
        pytest.skip("fixture missing")
    got = _run(tsan_exe, 3, 4, 1 | 2, [f])             # READBACK | SPLIT_GOPS: 1080p blocks through the pool, GOP pre-scan
    assert got["pictures"] == got["frames_out"] > 0 and got["pictures"] % 4 == 0 and len(set(got["hashes"])) == 1


def test_decoder_facade_producer_thread_under_tsan(tmp_path):
    """CH264VideoDecoder::open (the same-name shim over csrc/host/H264VideoDecoderB200.cpp): the front end runs on a producer thread, the
    callback on the calling thread, picture blocks go back to the pool from the consumer side — all of it under ThreadSanitizer against
    the engine test double; the callback protocol (frames in output order, final NULL + FILE_END) is checked by the client itself."""
    exe = str(tmp_path / "shim_tsan")
    srcs = [os.path.join(HERE, "shim_main_check.cpp"), os.path.join(HERE, "mock_engine.cpp")] + [os.path.join(HOST, f) for f in ("H264VideoDecoderB200.cpp", "h264_multi.cpp", "h264_front.cpp", "h264_slice.cpp", "h264_params.cpp")]
    r = subprocess.run(["g++", "-O1", "-g", "-fsanitize=thread", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", HOST, "-o", exe] + srcs + ["-pthread"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build not available: " + r.stderr[-300:])
    out = tmp_path / "bmp"
    out.mkdir()
    for name in ("synth_b_direct.first7.h264", "synth_mmco_lt_36.first9.h264", "HeavyHand_1080p.B_frames_cabac_tff.first7.h264"):
        f = os.path.join(HERE, "golden", name)
        if not os.path.exists(f):
            continue
        r = subprocess.run([exe, f, str(out)], capture_output=True, text=True, timeout=600)
        assert "ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
        assert r.returncode == 0 and "end_seen=1" in r.stdout, r.stdout[-500:] + r.stderr[-500:]
