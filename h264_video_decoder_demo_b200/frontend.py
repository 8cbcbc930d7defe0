"""ctypes access to the native host front end (include/h264_front_b200.h, built into libh264b2_host.so): the serial
entropy-decoding + derivation stage that turns an Annex-B H.264 byte stream into the per-picture structure-of-arrays
the CUDA engine consumes.  Pure host code; it produces no pixels."""
import ctypes as C
import os

from . import build as _build
from .engine import load_library

_LIB = None


class MultiStats(C.Structure):          # H264B2MultiStats (include/h264_multi_b200.h)
    _fields_ = [("seconds", C.c_double), ("parse_seconds", C.c_double), ("pictures", C.c_int64), ("frames_out", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("threads", C.c_int32), ("streams", C.c_int32),
                ("submits", C.c_int32), ("width_mbs", C.c_int32), ("height_mbs", C.c_int32), ("units", C.c_int32), ("reserved", C.c_int32)]


def lib():
    global _LIB
    if _LIB is None:
        load_library()                       # libh264b2.so first: the host library links against it
        if not os.path.exists(_build.HOST_LIB):
            _build.build_host()
        l = C.CDLL(_build.HOST_LIB)
        l.h264b2_front_write_container.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        l.h264b2_front_write_container.restype = C.c_int
        l.h264b2_multi_decode.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(MultiStats), C.c_char_p, C.c_size_t]
        l.h264b2_multi_decode.restype = C.c_int
        l.h264b2_front_write_container_range.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_size_t, C.c_size_t, C.c_int]
        l.h264b2_front_write_container_range.restype = C.c_int
        l.h264b2_front_gop_offsets.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
        l.h264b2_front_gop_offsets.restype = C.c_int
        _LIB = l
    return _LIB


def parse_to_container(h264_path: str, container_path: str, max_pictures: int = 0, begin: int = 0, end: int = 0, more_follows: bool = False) -> int:
    """Parse a byte stream (or its closed-GOP shard [begin, end)) and write its pictures (decoding order) + output order as a
    picture container (same format as oracle/ref_harness --replay, checksum fields 0).  Returns 0 or a negative error code."""
    return lib().h264b2_front_write_container_range(os.fsencode(h264_path), os.fsencode(container_path), int(max_pictures), int(begin), int(end), 1 if more_follows else 0)


def multi_decode(paths, device=0, threads=1, readback=True, hashes=True, split_gops=False):
    """Decode many Annex-B streams on one GPU (parser thread pool -> batched submits, see include/h264_multi_b200.h).
    Returns (stats dict, per-stream hash chains or None).  Raises RuntimeError on failure — there is no CPU fallback."""
    n = len(paths)
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    st = MultiStats()
    hs = (C.c_uint64 * n)() if hashes else None
    err = C.create_string_buffer(512)
    r = lib().h264b2_multi_decode(device, n, arr, int(threads), (1 if readback else 0) | (2 if split_gops else 0), hs, C.byref(st), err, 512)
    if r != 0:
        raise RuntimeError(f"h264b2_multi_decode failed ({r}): {err.value.decode(errors='replace')}")
    stats = {f: getattr(st, f) for f, _ in MultiStats._fields_}
    return stats, ([int(x) for x in hs] if hashes else None)


def hash_chain(sums):
    """The per-stream chain h264b2_multi_decode reports, from a list of per-frame checksums in output order."""
    h = 0
    for s in sums:
        h = (h * 0x100000001B3 + s) & 0xFFFFFFFFFFFFFFFF
    return h


def gop_offsets(data: bytes):
    """Byte offsets of the closed GOPs of an Annex-B stream (first parameter-set NAL in front of every IDR picture)."""
    n = lib().h264b2_front_gop_offsets(data, len(data), None, 0)
    arr = (C.c_size_t * max(n, 1))()
    lib().h264b2_front_gop_offsets(data, len(data), arr, n)
    return [int(arr[i]) for i in range(n)]
