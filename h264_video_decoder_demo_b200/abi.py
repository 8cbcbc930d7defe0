"""ctypes / numpy mirror of include/h264_recon_b200.h (the C ABI of the B200 reconstruction engine).

Every layout here must match the header byte for byte; tests/test_abi.py checks sizes
against the compiled library.
"""
import ctypes as C
import numpy as np

ABI_VERSION = 2

# macroblock classes (H264B2_MB_*)
MB_NA, MB_I4x4, MB_I8x8, MB_I16x16, MB_IPCM, MB_INTER = range(6)
MBF_FIELD, MBF_T8x8, MBF_SPSI, MBF_CIP_UNAVAIL = 1, 2, 4, 8

CM_LUMA_DC = 1 << 16
CM_CHROMA_DC = 1 << 17
CM_PCM = 1 << 26


def CM_LUMA(b):
    return 1 << b


def CM_CB(b):
    return 1 << (18 + b)


def CM_CR(b):
    return 1 << (22 + b)


MB_INFO_DT = np.dtype([
    ("mb_class", "u1"), ("flags", "u1"), ("pred16_chroma", "u1"), ("qpy", "i1"),
    ("slice_number", "<u2"), ("nnz_mask", "<u2"),
    ("filter_offset_a", "i1"), ("filter_offset_b", "i1"), ("deblock_idc", "u1"), ("reserved", "u1"),
    ("coef_mask", "<u4"),
])
MB_MOTION_DT = np.dtype([
    ("mv", "<i2", (2, 16, 2)), ("ref_surf", "i1", (2, 4)), ("ref_ident", "i1", (2, 4)), ("wt_idx", "<u2", (4,)),
])
WEIGHT_DT = np.dtype([
    ("mode", "<i2"), ("logwd", "<i2", (3,)), ("w0", "<i2", (3,)), ("w1", "<i2", (3,)), ("o0", "<i2", (3,)), ("o1", "<i2", (3,)),
])
assert MB_INFO_DT.itemsize == 16 and MB_MOTION_DT.itemsize == 152 and WEIGHT_DT.itemsize == 32


class PicParams(C.Structure):
    _fields_ = [
        ("width_mbs", C.c_int32), ("height_mbs", C.c_int32), ("mbaff_frame_flag", C.c_int32),
        ("chroma_qp_offset", C.c_int32 * 2), ("dst_surface", C.c_int32), ("clear_surface", C.c_int32),
        ("has_inter", C.c_int32), ("deblock_enable", C.c_int32), ("deblock_stop_mb", C.c_int32),
        ("n_weights", C.c_int32), ("n_coefs", C.c_uint32), ("custom_scaling", C.c_int32), ("packed", C.c_int32),
        ("mb_info", C.c_void_p), ("intra_modes", C.c_void_p), ("coef_offset", C.c_void_p), ("motion", C.c_void_p),
        ("weights", C.c_void_p), ("coefs", C.c_void_p), ("level_scale4", C.c_void_p), ("level_scale8", C.c_void_p),
    ]


assert C.sizeof(PicParams) == 56 + 64

_K = np.uint64(0x9E3779B97F4A7C15)


def checksum(buf) -> int:
    """Positional 64-bit checksum used by the engine, the harness and the oracle:
    sum over LE u32 words w_i of (w_i + 1) * ((2 i + 1) * K) mod 2^64."""
    w = np.frombuffer(buf, dtype="<u4").astype(np.uint64)
    i = np.arange(w.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(((w + np.uint64(1)) * ((np.uint64(2) * i + np.uint64(1)) * _K)).sum(dtype=np.uint64))
