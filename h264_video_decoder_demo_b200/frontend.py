"""ctypes access to the native host front end (include/h264_front_b200.h, built into libh264b2_host.so): the serial
entropy-decoding + derivation stage that turns an Annex-B H.264 byte stream into the per-picture structure-of-arrays
the CUDA engine consumes.  Pure host code; it produces no pixels."""
import ctypes as C
import os

from . import build as _build
from .engine import load_library

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        load_library()                       # libh264b2.so first: the host library links against it
        if not os.path.exists(_build.HOST_LIB):
            _build.build_host()
        l = C.CDLL(_build.HOST_LIB)
        l.h264b2_front_write_container.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        l.h264b2_front_write_container.restype = C.c_int
        _LIB = l
    return _LIB


def parse_to_container(h264_path: str, container_path: str, max_pictures: int = 0) -> int:
    """Parse a whole byte stream and write its pictures (decoding order) + output order as a picture container
    (same format as oracle/ref_harness --replay, checksum fields 0).  Returns 0 or a negative error code."""
    return lib().h264b2_front_write_container(os.fsencode(h264_path), os.fsencode(container_path), int(max_pictures))
