"""ctypes binding of the CUDA engine's C ABI (include/h264_recon_b200.h -> libh264b2.so).

This is the only way the Python side reaches the product: there is no fallback.  If the shared library
is missing it is built with nvcc (build.py); if that fails, or no CUDA device exists, every entry point
raises.
"""
import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import build as _build
from .abi import PicParams
from . import replay as _replay

_LIB = None

_PROTOS = {
    # name: (restype, argtypes)
    "h264b2_abi_version": (C.c_int, []),
    "h264b2_last_error": (C.c_char_p, []),
    "h264b2_checksum_host": (C.c_uint64, [C.c_void_p, C.c_size_t]),
    "h264b2_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "h264b2_destroy": (C.c_int, [C.c_void_p]),
    "h264b2_submit": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(PicParams)]),
    "h264b2_submit_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(PicParams)]),
    "h264b2_read_picture": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "h264b2_read_pictures_async": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]),
    "h264b2_read_picture_bgr24": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "h264b2_write_picture": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "h264b2_checksum_picture": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "h264b2_checksum_pictures": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "h264b2_surface_ptr": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "h264b2_dev_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "h264b2_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "h264b2_dev_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "h264b2_dev_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "h264b2_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "h264b2_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "h264b2_sync": (C.c_int, [C.c_void_p]),
    "h264b2_set_lookahead": (C.c_int, [C.c_void_p, C.c_int]),
    "h264b2_debug_intra_tables": (C.c_int, [C.c_int, C.c_void_p]),
    "h264b2_pack_coefs_bound": (C.c_size_t, [C.c_uint32]),
    "h264b2_pack_coefs": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "h264b2_unpack_coefs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "h264b2_pack_motion": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "h264b2_unpack_motion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "h264b2_timer_start": (C.c_int, [C.c_void_p]),
    "h264b2_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "h264b2_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
}
ABI_SYMBOLS = sorted(_PROTOS)
KERNEL_CLASSES = ("clear", "inter", "intra", "bs", "deblock", "residual")


class EngineError(RuntimeError):
    pass


def library_path() -> str:
    """The engine library this process loads (H264B2_LIB overrides the in-tree build for A/B runs)."""
    return os.environ.get("H264B2_LIB", _build.LIB)


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """Load libh264b2.so (building it first when absent).  Needs no GPU: only symbol resolution."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_build.LIB):
            if not build_if_missing:
                raise EngineError(f"{_build.LIB} is missing and there is no CPU fallback; run __graft_entry__.build()")
            _build.build()
        lib = C.CDLL(os.environ.get("H264B2_LIB", _build.LIB))      # H264B2_LIB: A/B a differently compiled engine
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)       # raises AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


def pack_coefs(dense: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Packed coefficient blob (h264b2_pack_coefs) of a dense int16 level array; a host function, no GPU needed.
    `out`: optional uint8 buffer (16-byte aligned, e.g. page-locked) to pack into; returns the blob as a view of it."""
    lib = load_library()
    dense = np.ascontiguousarray(dense, dtype=np.int16)
    bound = lib.h264b2_pack_coefs_bound(dense.size)
    if out is None:
        raw = np.empty(bound + 16, np.uint8)
        o = (-raw.ctypes.data) % 16
        out = raw[o:o + bound]
    n = C.c_size_t()
    rc = lib.h264b2_pack_coefs(dense.ctypes.data if dense.size else None, dense.size, out.ctypes.data, out.size, C.byref(n))
    if rc != 0:
        raise EngineError(f"h264b2_pack_coefs: {lib.h264b2_last_error().decode(errors='replace')}")
    return out[:n.value]


def pack_motion(motion: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Packed blob (h264b2_pack_motion) of a picture's H264B2MbMotion records (abi.MB_MOTION_DT)."""
    lib = load_library()
    motion = np.ascontiguousarray(motion)
    n = motion.size
    bound = lib.h264b2_pack_coefs_bound(n * 76)
    if out is None:
        raw = np.empty(bound + 16, np.uint8)
        o = (-raw.ctypes.data) % 16
        out = raw[o:o + bound]
    nb = C.c_size_t()
    if lib.h264b2_pack_motion(motion.ctypes.data if n else None, n, out.ctypes.data, out.size, C.byref(nb)) != 0:
        raise EngineError(f"h264b2_pack_motion: {lib.h264b2_last_error().decode(errors='replace')}")
    return out[:nb.value]


def unpack_motion(blob: np.ndarray, n_mbs: int) -> np.ndarray:
    from .abi import MB_MOTION_DT
    lib = load_library()
    m = np.zeros(n_mbs, MB_MOTION_DT)
    if lib.h264b2_unpack_motion(blob.ctypes.data, m.ctypes.data if n_mbs else None, n_mbs) != 0:
        raise EngineError(f"h264b2_unpack_motion: {lib.h264b2_last_error().decode(errors='replace')}")
    return m


def unpack_coefs(blob: np.ndarray, n_coefs: int) -> np.ndarray:
    lib = load_library()
    dense = np.empty(n_coefs, np.int16)
    if lib.h264b2_unpack_coefs(blob.ctypes.data, dense.ctypes.data if n_coefs else None, n_coefs) != 0:
        raise EngineError(f"h264b2_unpack_coefs: {lib.h264b2_last_error().decode(errors='replace')}")
    return dense


class Engine:
    """One per-GPU reconstruction context (H264B2Context)."""

    def __init__(self, device: int, n_streams: int, width_mbs: int, height_mbs: int, surfaces_per_stream: int = 17):
        self.lib = load_library()
        self.n_streams, self.wmb, self.hmb, self.spp = n_streams, width_mbs, height_mbs, surfaces_per_stream
        self.frame_bytes = width_mbs * height_mbs * 384
        self._ctx = C.c_void_p()
        self._ck(self.lib.h264b2_create(C.byref(self._ctx), device, n_streams, surfaces_per_stream, width_mbs, height_mbs))

    def _ck(self, rc: int):
        if rc != 0:
            raise EngineError(f"h264b2 error {rc}: {self.lib.h264b2_last_error().decode(errors='replace')}")

    def close(self):
        if self._ctx:
            self.lib.h264b2_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- submit
    @staticmethod
    def _arrays(stream_ids: Sequence[int], params: Sequence[PicParams]):
        n = len(params)
        sid = (C.c_int32 * n)(*stream_ids)
        arr = (PicParams * n)(*params)
        return n, sid, arr

    def submit(self, stream_ids, params):
        n, sid, arr = self._arrays(stream_ids, params)
        self._ck(self.lib.h264b2_submit(self._ctx, n, sid, arr))

    def submit_device(self, stream_ids, params):
        n, sid, arr = self._arrays(stream_ids, params)
        self._ck(self.lib.h264b2_submit_device(self._ctx, n, sid, arr))

    def submit_prepared(self, prepared):
        """prepared = (n, c_int32 array, PicParams array) built once by prepare(); avoids per-call marshalling."""
        self._ck(self.lib.h264b2_submit_device(self._ctx, prepared[0], prepared[1], prepared[2]))

    def submit_prepared_host(self, prepared):
        self._ck(self.lib.h264b2_submit(self._ctx, prepared[0], prepared[1], prepared[2]))

    prepare = _arrays

    # ---- pictures
    def read_picture(self, stream_id: int, surface: int) -> np.ndarray:
        out = np.empty(self.frame_bytes, dtype=np.uint8)
        self._ck(self.lib.h264b2_read_picture(self._ctx, stream_id, surface, out.ctypes.data))
        return out

    def read_picture_bgr24(self, stream_id: int, surface: int, width_bytes: Optional[int] = None, flip_lines: bool = False) -> np.ndarray:
        width_bytes = width_bytes or self.wmb * 48
        out = np.empty(width_bytes * self.hmb * 16, dtype=np.uint8)
        self._ck(self.lib.h264b2_read_picture_bgr24(self._ctx, stream_id, surface, out.ctypes.data, width_bytes, 1 if flip_lines else 0))
        return out

    def write_picture(self, stream_id: int, surface: int, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.size == self.frame_bytes
        self._ck(self.lib.h264b2_write_picture(self._ctx, stream_id, surface, data.ctypes.data))

    def read_pictures_async(self, stream_ids, surfaces, host_ptrs):
        n = len(stream_ids)
        self._ck(self.lib.h264b2_read_pictures_async(self._ctx, n, (C.c_int32 * n)(*stream_ids), (C.c_int32 * n)(*surfaces),
                                                     (C.c_void_p * n)(*host_ptrs)))

    def checksum(self, stream_id: int, surface: int) -> int:
        v = C.c_uint64()
        self._ck(self.lib.h264b2_checksum_picture(self._ctx, stream_id, surface, C.byref(v)))
        return v.value

    def checksums(self, stream_ids, surfaces) -> List[int]:
        n = len(stream_ids)
        out = (C.c_uint64 * n)()
        self._ck(self.lib.h264b2_checksum_pictures(self._ctx, n, (C.c_int32 * n)(*stream_ids), (C.c_int32 * n)(*surfaces), out))
        return list(out)

    # ---- memory
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.lib.h264b2_dev_alloc(self._ctx, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, ptr: int):
        self._ck(self.lib.h264b2_dev_free(self._ctx, ptr))

    def dev_upload(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._ck(self.lib.h264b2_dev_upload(self._ctx, dst, src.ctypes.data, src.nbytes))

    def dev_copy(self, dst: int, src: int, nbytes: int):
        self._ck(self.lib.h264b2_dev_copy(self._ctx, dst, src, nbytes))

    def host_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.lib.h264b2_host_alloc(self._ctx, nbytes, C.byref(p)))
        return p.value

    def host_free(self, ptr: int):
        self._ck(self.lib.h264b2_host_free(self._ctx, ptr))

    def pinned_array(self, nbytes: int) -> np.ndarray:
        """uint8 numpy view over freshly allocated page-locked memory (freed with the engine's process)."""
        ptr = self.host_alloc(nbytes)
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr))

    def set_lookahead(self, on: bool):
        self._ck(self.lib.h264b2_set_lookahead(self._ctx, 1 if on else 0))

    def sync(self):
        self._ck(self.lib.h264b2_sync(self._ctx))

    # ---- timing
    def timer_start(self):
        self._ck(self.lib.h264b2_timer_start(self._ctx))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.h264b2_timer_stop(self._ctx, C.byref(ms)))
        return ms.value

    def kernel_times(self) -> Dict[str, Dict[str, float]]:
        ms = (C.c_float * 6)()
        n = (C.c_int64 * 6)()
        self._ck(self.lib.h264b2_kernel_times(self._ctx, ms, n))
        return {k: {"ms": float(ms[i]), "launches": int(n[i])} for i, k in enumerate(KERNEL_CLASSES)}


_ARRAY_NAMES = ("mb_info", "intra_modes", "coef_offset", "motion", "weights", "coefs", "level_scale4", "level_scale8")


class ResidentStream:
    """A pre-parsed stream (replay container) made resident in HBM: one device blob per picture.

    params[i] is the PicParams of picture i (decoding order) with DEVICE pointers, ready for
    Engine.submit_device().  clone() makes an independent device copy (separate HBM bytes)."""

    def __init__(self, eng: Engine, rp: "_replay.Replay", _blobs=None):
        self.eng, self.rp = eng, rp
        self.blobs: List[int] = []
        self.blob_bytes: List[int] = []
        self.params: List[PicParams] = []
        for i, pic in enumerate(rp.pictures):
            offs, total = {}, 0
            for name in _ARRAY_NAMES:
                a = getattr(pic, name)
                if a is None or a.size == 0:
                    continue
                offs[name] = total
                total += (a.nbytes + 255) & ~255
            if _blobs is None:
                host = np.zeros(total, dtype=np.uint8)
                for name, o in offs.items():
                    a = getattr(pic, name)
                    host[o:o + a.nbytes] = np.frombuffer(a.tobytes(), dtype=np.uint8)
                d = eng.dev_alloc(total)
                eng.dev_upload(d, host)
            else:
                d = eng.dev_alloc(total)
                eng.dev_copy(d, _blobs[i], total)
            self.blobs.append(d)
            self.blob_bytes.append(total)
            self.params.append(_replay.pic_params(rp, pic, ptrs={k: d + o for k, o in offs.items()}))

    def clone(self) -> "ResidentStream":
        return ResidentStream(self.eng, self.rp, _blobs=self.blobs)

    def free(self):
        for d in self.blobs:
            self.eng.dev_free(d)
        self.blobs = []
