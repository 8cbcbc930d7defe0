"""The host front end under AddressSanitizer + UndefinedBehaviorSanitizer: random-syntax streams, streams of random bytes behind valid
headers, and corrupted / truncated prefixes of the committed fixtures must parse (or fail cleanly) without a single report.  The parser
reads attacker-controlled bytes and mirrors several of the reference's failure modes on purpose — those must stay inside
its own arrays.  CPU only; builds tools/h264b2_parse.cpp + csrc/host with -fsanitize=address,undefined."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
HOST = os.path.join(ROOT, "h264_video_decoder_demo_b200", "csrc", "host")


@pytest.fixture(scope="module")
def asan_parser(tmp_path_factory):
    d = tmp_path_factory.mktemp("asan")
    stub = d / "stubs.c"      # the packers live in the CUDA engine's library; the parser only calls them in packed mode (not used here)
    stub.write_text("#include <stddef.h>\nlong h264b2_pack_coefs(void){return -1;} size_t h264b2_pack_coefs_bound(size_t n){return n*4;} long h264b2_pack_motion(void){return -1;}\n")
    exe = str(d / "parse_asan")
    r = subprocess.run(["gcc", "-c", str(stub), "-o", str(d / "stubs.o")], capture_output=True, text=True)
    if r.returncode == 0:
        r = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", HOST, "-o", exe,
                            os.path.join(ROOT, "tools", "h264b2_parse.cpp")] + [os.path.join(HOST, f) for f in ("h264_front.cpp", "h264_slice.cpp", "h264_params.cpp")] + [str(d / "stubs.o"), "-pthread"],
                           capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available: " + r.stderr[-300:])
    return exe, d


def _run(exe, d, data):
    src = str(d / "s.h264")
    open(src, "wb").write(data)
    r = subprocess.run([exe, src, str(d / "o.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode >= 0, f"killed by signal {-r.returncode}: {r.stderr[-800:]}"
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-1500:]


def test_generated_and_noise_streams_are_clean_under_sanitizers(asan_parser):
    import h264_writer
    exe, d = asan_parser
    seed0 = int.from_bytes(os.urandom(2), "little")
    for k in range(12):
        for kind in ("I", "P", "B"):
            _run(exe, d, h264_writer.random_cabac_stream(seed0 + k, kind=kind, wmb=6, hmb=5, n_pics=3, nbytes=400, t8x8=bool(k & 1)))
    cfgs = [dict(), dict(t8x8=True, weighted=True), dict(bframes=True, bipred_idc=1), dict(bframes=True, bipred_idc=2), dict(mmco=True, n_pics=8, n_refs=4, max_slices=2), dict(fn_gaps=True)]
    for k, cfg in enumerate(cfgs * 2):
        _run(exe, d, h264_writer.Stream(seed=seed0 + k, **cfg).build())


def test_corrupted_fixture_prefixes_are_clean_under_sanitizers(asan_parser):
    exe, d = asan_parser
    files = sorted(glob.glob(os.path.join(HERE, "golden", "*.h264")))
    if not files:
        pytest.skip("no Annex-B fixtures")
    rng = np.random.default_rng(int.from_bytes(os.urandom(2), "little"))
    for f in files:
        raw = open(f, "rb").read()
        big = len(raw) > 100000                       # the 1080p prefixes (incl. the MBAFF stream) take seconds each under ASan: one variant
        for k in range(1 if big else 2):
            b = bytearray(raw[:max(200, len(raw) // (k + 1))])
            for _ in range(20):
                b[int(rng.integers(50, len(b)))] ^= int(rng.integers(1, 256))
            _run(exe, d, bytes(b))
