import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return lib.cuInit(0) == 0 and lib.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
FULL_DIR = os.path.join(ROOT, "oracle", "_ref", "replay")


def golden_files(kind="bundled"):
    """kind: "bundled" = 1080p prefixes of the reference's own streams, "synth" = small random-syntax streams
    (tests/h264_writer.py, decoded by the reference), "all"."""
    fs = sorted(os.path.join(GOLDEN_DIR, f) for f in os.listdir(GOLDEN_DIR) if f.endswith(".rp.xz"))
    if kind == "all":
        return fs
    return [f for f in fs if os.path.basename(f).startswith("synth_") == (kind == "synth")]


def full_files():
    if not os.path.isdir(FULL_DIR):
        return []
    return sorted(os.path.join(FULL_DIR, f) for f in os.listdir(FULL_DIR) if f.endswith(".bin.xz"))
