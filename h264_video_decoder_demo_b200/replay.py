"""Pre-parsed picture containers ("replay files").

A replay file holds, for every picture of a stream in DECODING order, the per-picture
structure-of-arrays the engine consumes (include/h264_recon_b200.h) -- i.e. the output of the
serial host stage (entropy decode + derivations) -- plus the output order and checksums of the
reference decoder's pictures.  Files are produced by oracle/ref_harness.cpp from the unmodified
reference (and, once the native host parser covers a stream, by the parser itself); this module
only reads them.
"""
import lzma
import os
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .abi import MB_INFO_DT, MB_MOTION_DT, WEIGHT_DT, PicParams

_FILE_HDR = struct.Struct("<8s8I")
_PIC_HDR = struct.Struct("<14i2I2Q")


@dataclass
class Picture:
    decode_idx: int
    dst_surface: int
    clear_surface: int
    has_inter: int
    deblock_enable: int
    deblock_stop_mb: int
    mbaff: int
    cqp: tuple
    slice_type: int
    poc: int
    n_na: int
    nal_ref_idc: int
    sum_pre: int
    sum_post: int
    mb_info: np.ndarray
    intra_modes: np.ndarray
    coef_offset: np.ndarray
    motion: Optional[np.ndarray]
    weights: np.ndarray
    coefs: np.ndarray
    level_scale4: Optional[np.ndarray] = None
    level_scale8: Optional[np.ndarray] = None

    def nbytes(self) -> int:
        n = self.mb_info.nbytes + self.intra_modes.nbytes + self.coef_offset.nbytes + self.weights.nbytes + self.coefs.nbytes
        if self.motion is not None:
            n += self.motion.nbytes
        return n


@dataclass
class Replay:
    path: str
    width_mbs: int
    height_mbs: int
    pictures: List[Picture] = field(default_factory=list)
    out_order: List[int] = field(default_factory=list)     # decode_idx per output frame
    out_sums: List[int] = field(default_factory=list)      # reference checksum per output frame

    @property
    def n_mbs(self):
        return self.width_mbs * self.height_mbs

    @property
    def frame_bytes(self):
        return self.n_mbs * 384


def _read_all(path: str) -> bytes:
    if path.endswith(".xz"):
        with lzma.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def load_replay(path: str, max_pictures: Optional[int] = None) -> Replay:
    return parse_replay(_read_all(path), path, max_pictures)


def read_replay_bytes(path: str) -> bytes:
    """Raw (decompressed) container bytes, e.g. to place them in page-locked memory before parsing."""
    return _read_all(path)


def parse_replay(buf, path: str = "<buffer>", max_pictures: Optional[int] = None) -> Replay:
    """Parse a container held in any buffer object; the numpy arrays are VIEWS into `buf` (so a pinned
    buffer yields pinned, per-picture contiguous arrays that h264b2_submit can DMA in one transfer)."""
    magic, version, wmb, hmb, n_pics, n_out, hdr_bytes, pichdr_bytes, _ = _FILE_HDR.unpack_from(buf, 0)
    if magic != b"H264B2RP" or version != 1:
        raise ValueError(f"{path}: not a replay file")
    assert hdr_bytes == _FILE_HDR.size and pichdr_bytes == _PIC_HDR.size
    rp = Replay(path, wmb, hmb)
    nmb = wmb * hmb
    off = hdr_bytes
    for _ in range(n_pics):
        (didx, dst, clr, has_inter, dbk, stop, mbaff, cqp0, cqp1, n_w, custom, stype, poc, n_na,
         n_coefs, nri, sum_pre, sum_post) = _PIC_HDR.unpack_from(buf, off)
        off += pichdr_bytes

        def take(dt, n):
            nonlocal off
            a = np.frombuffer(buf, dtype=dt, count=n, offset=off)
            off += a.nbytes
            return a

        mb_info = take(MB_INFO_DT, nmb)
        modes = take("<u8", nmb)
        coff = take("<u4", nmb)
        motion = take(MB_MOTION_DT, nmb) if has_inter else None
        weights = take(WEIGHT_DT, n_w)
        coefs = take("<i2", n_coefs)
        ls4 = ls8 = None
        if custom:
            ls4 = take("<i2", 2 * 2 * 6 * 16)
            ls8 = take("<i2", 2 * 2 * 6 * 64)
        if max_pictures is None or len(rp.pictures) < max_pictures:
            rp.pictures.append(Picture(didx, dst, clr, has_inter, dbk, stop, mbaff, (cqp0, cqp1), stype, poc, n_na, nri,
                                       sum_pre, sum_post, mb_info, modes, coff, motion, weights, coefs, ls4, ls8))
    out = np.frombuffer(buf, dtype=np.dtype([("idx", "<i4"), ("pad", "<i4"), ("sum", "<u8")]), count=n_out, offset=off)
    rp.out_order = [int(x) for x in out["idx"]]
    rp.out_sums = [int(x) for x in out["sum"]]
    return rp


def pic_params(rp: Replay, pic: Picture, ptrs=None, packed_blob=None, packed_motion=None) -> PicParams:
    """Fill a PicParams for `pic`.  `ptrs` maps array name -> integer address (host or device);
    when None the numpy arrays' own host addresses are used (keep `pic` alive while in use).
    `packed_blob`: uint8 array from engine.pack_coefs(pic.coefs) -> the picture travels with packed coefficients;
    `packed_motion`: uint8 array from engine.pack_motion(pic.motion) -> ... with packed motion records (h264b2_submit only)."""
    p = PicParams()
    p.width_mbs, p.height_mbs, p.mbaff_frame_flag = rp.width_mbs, rp.height_mbs, pic.mbaff
    p.chroma_qp_offset[0], p.chroma_qp_offset[1] = pic.cqp
    p.dst_surface, p.clear_surface, p.has_inter = pic.dst_surface, pic.clear_surface, pic.has_inter
    p.deblock_enable, p.deblock_stop_mb = pic.deblock_enable, pic.deblock_stop_mb
    p.n_weights, p.n_coefs = len(pic.weights), len(pic.coefs)
    p.custom_scaling = 1 if pic.level_scale4 is not None else 0

    def addr(name):
        if ptrs is not None:
            return ptrs.get(name) or None
        a = getattr(pic, name)
        return a.ctypes.data if a is not None and a.size else None

    for name in ("mb_info", "intra_modes", "coef_offset", "motion", "weights", "coefs", "level_scale4", "level_scale8"):
        setattr(p, name, addr(name))
    if packed_blob is not None and len(pic.coefs):
        p.packed, p.coefs = p.packed | 1, packed_blob.ctypes.data
    if packed_motion is not None and pic.has_inter and pic.motion is not None:
        p.packed, p.motion = p.packed | 2, packed_motion.ctypes.data
    return p


def default_replay_dir() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "replay")


def save_replay(rp: Replay, path: str, n_pictures: Optional[int] = None, preset: int = 6) -> None:
    """Write the first n_pictures (decoding order) of `rp` as a replay container (xz if path ends with .xz)."""
    pics = rp.pictures[:n_pictures] if n_pictures is not None else rp.pictures
    keep = {p.decode_idx for p in pics}
    outs = [(i, s) for i, s in zip(rp.out_order, rp.out_sums) if i in keep]
    parts = [_FILE_HDR.pack(b"H264B2RP", 1, rp.width_mbs, rp.height_mbs, len(pics), len(outs), _FILE_HDR.size, _PIC_HDR.size, 0)]
    for p in pics:
        parts.append(_PIC_HDR.pack(p.decode_idx, p.dst_surface, p.clear_surface, p.has_inter, p.deblock_enable, p.deblock_stop_mb,
                                   p.mbaff, p.cqp[0], p.cqp[1], len(p.weights), 1 if p.level_scale4 is not None else 0,
                                   p.slice_type, p.poc, p.n_na, len(p.coefs), p.nal_ref_idc, p.sum_pre, p.sum_post))
        parts += [p.mb_info.tobytes(), p.intra_modes.tobytes(), p.coef_offset.tobytes()]
        if p.has_inter:
            parts.append(p.motion.tobytes())
        parts += [p.weights.tobytes(), p.coefs.tobytes()]
        if p.level_scale4 is not None:
            parts += [p.level_scale4.tobytes(), p.level_scale8.tobytes()]
    out = np.zeros(len(outs), dtype=np.dtype([("idx", "<i4"), ("pad", "<i4"), ("sum", "<u8")]))
    out["idx"] = [i for i, _ in outs]
    out["sum"] = [s for _, s in outs]
    parts.append(out.tobytes())
    blob = b"".join(parts)
    if path.endswith(".xz"):
        with lzma.open(path, "wb", preset=preset) as f:
            f.write(blob)
    else:
        with open(path, "wb") as f:
            f.write(blob)
