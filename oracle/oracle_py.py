"""ctypes loader for the CPU oracle (oracle/liboracle_recon.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_recon.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_reconstruct_picture.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int]
        _LIB.oracle_reconstruct_picture.restype = C.c_int
        _LIB.oracle_checksum.argtypes = [C.c_void_p, C.c_size_t]
        _LIB.oracle_checksum.restype = C.c_uint64
        _LIB.oracle_convert_bgr24.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _LIB.oracle_convert_bgr24.restype = C.c_int
    return _LIB


STAGE_RECON, STAGE_DEBLOCK = 1, 2


class OracleDPB:
    """n_surfaces host I420 surfaces + the oracle's reconstruct call."""

    def __init__(self, width_mbs, height_mbs, n_surfaces=17):
        self.frame_bytes = width_mbs * height_mbs * 384
        # one contiguous allocation + zero tail padding, exactly like the engine's DPB: a bottom-field view
        # clamps x to 2W-1 (Q4) and can read up to one chroma row past the end of a surface, i.e. into the
        # next surface (undefined behaviour in the reference; defined as "what follows in the DPB" here)
        self._store = np.zeros(n_surfaces * self.frame_bytes + width_mbs * 64, dtype=np.uint8)
        self.surfaces = [self._store[i * self.frame_bytes:(i + 1) * self.frame_bytes] for i in range(n_surfaces)]
        self._ptrs = (C.c_void_p * n_surfaces)(*[s.ctypes.data for s in self.surfaces])
        self.n = n_surfaces

    def reconstruct(self, params, stages=STAGE_RECON | STAGE_DEBLOCK):
        r = lib().oracle_reconstruct_picture(C.addressof(params), self._ptrs, self.n, stages)
        if r != 0:
            raise RuntimeError(f"oracle_reconstruct_picture failed: {r}")

    def checksum(self, slot):
        s = self.surfaces[slot]
        return int(lib().oracle_checksum(s.ctypes.data, s.size))


def convert_bgr24(i420: np.ndarray, W: int, H: int, flip: bool = False) -> np.ndarray:
    out = np.zeros(W * 3 * H, dtype=np.uint8)
    src = np.ascontiguousarray(i420, dtype=np.uint8)
    if lib().oracle_convert_bgr24(src.ctypes.data, W, H, out.ctypes.data, W * 3, 1 if flip else 0) != 0:
        raise RuntimeError("oracle_convert_bgr24 failed")
    return out
