"""Python mirror of the reference's decoder interface (H264VideoDecoder.h:22-43) over the C++ host facade
(libh264b2_host.so -> CH264VideoDecoderB200 -> the CUDA engine).  Same names and conventions as the reference:
methods return int (0 ok), open() blocks and calls back synchronously, once per output picture in display order
and finally with outPicture None and errorCode H264_DECODE_ERROR_CODE_FILE_END; a non-zero callback return stops
decoding.  The picture's planes are only valid during the callback.
"""
import ctypes as C
import lzma
import os
import tempfile

import numpy as np

from . import build as _build
from .engine import load_library

H264_DECODE_ERROR_CODE_NO = 0
H264_DECODE_ERROR_CODE_FILE_END = 1


class _PicBase(C.Structure):
    _fields_ = [("m_pic_buff_luma", C.POINTER(C.c_uint8)), ("m_pic_buff_cb", C.POINTER(C.c_uint8)), ("m_pic_buff_cr", C.POINTER(C.c_uint8)),
                ("PicWidthInSamplesL", C.c_int32), ("PicHeightInSamplesL", C.c_int32), ("PicWidthInSamplesC", C.c_int32), ("PicHeightInSamplesC", C.c_int32),
                ("PicOrderCnt", C.c_int32), ("m_PicNumCnt", C.c_int32), ("slice_type", C.c_int32), ("MbaffFrameFlag", C.c_int32),
                ("profile_idc", C.c_int32), ("level_idc", C.c_int32), ("entropy_coding_mode_flag", C.c_int32), ("fps", C.c_float)]


class _Pic(C.Structure):
    _fields_ = [("m_picture_frame", _PicBase)]


_CB = C.CFUNCTYPE(C.c_int, C.POINTER(_Pic), C.c_void_p, C.c_int)
_HOST = None


def _host():
    global _HOST
    if _HOST is None:
        load_library()                      # libh264b2.so first (the facade links against it)
        if not os.path.exists(_build.HOST_LIB):
            _build.build_host()
        lib = C.CDLL(_build.HOST_LIB)
        lib.h264b2_decoder_create.restype = C.c_void_p
        lib.h264b2_decoder_destroy.argtypes = [C.c_void_p]
        lib.h264b2_decoder_set_callback.argtypes = [C.c_void_p, _CB, C.c_void_p]
        lib.h264b2_decoder_set_device.argtypes = [C.c_void_p, C.c_int]
        lib.h264b2_decoder_open.argtypes = [C.c_void_p, C.c_char_p]
        lib.h264b2_decoder_last_error.argtypes = [C.c_void_p]
        lib.h264b2_decoder_last_error.restype = C.c_char_p
        _HOST = lib
    return _HOST


class OutPicture:
    """What the callback receives: the fields consumers of the reference read from CH264Picture::m_picture_frame."""

    def __init__(self, p: _PicBase):
        self.PicWidthInSamplesL, self.PicHeightInSamplesL = p.PicWidthInSamplesL, p.PicHeightInSamplesL
        self.PicWidthInSamplesC, self.PicHeightInSamplesC = p.PicWidthInSamplesC, p.PicHeightInSamplesC
        self.PicOrderCnt, self.m_PicNumCnt, self.slice_type, self.MbaffFrameFlag = p.PicOrderCnt, p.m_PicNumCnt, p.slice_type, p.MbaffFrameFlag
        n = self.PicWidthInSamplesL * self.PicHeightInSamplesL * 3 // 2
        self.i420 = np.ctypeslib.as_array(p.m_pic_buff_luma, shape=(n,))    # Y|Cb|Cr contiguous; valid during the callback only


class H264VideoDecoder:
    def __init__(self):
        self._lib = _host()
        self._h = C.c_void_p(self._lib.h264b2_decoder_create())
        self._cb = None
        self._keep = None

    def init(self):
        return 0

    def unInit(self):
        return 0

    def set_output_frame_callback_functuin(self, output_frame_callback, userData=None):   # sic
        def tramp(pic, _user, err):
            try:
                return int(output_frame_callback(OutPicture(pic.contents.m_picture_frame) if pic else None, userData, err) or 0)
            except Exception:            # an exception must not unwind through C
                import traceback
                traceback.print_exc()
                return -1
        self._keep = _CB(tramp)
        return self._lib.h264b2_decoder_set_callback(self._h, self._keep, None)

    def set_device(self, device):
        return self._lib.h264b2_decoder_set_device(self._h, device)

    def open(self, url):
        tmp = None
        try:
            if url.endswith(".xz"):     # containers are shipped compressed; the C++ facade reads them raw
                tmp = tempfile.NamedTemporaryFile(suffix=".bin", delete=False)
                with lzma.open(url, "rb") as f:
                    tmp.write(f.read())
                tmp.close()
                url = tmp.name
            return self._lib.h264b2_decoder_open(self._h, os.fsencode(url))
        finally:
            if tmp is not None:
                os.unlink(tmp.name)

    def last_error(self):
        return self._lib.h264b2_decoder_last_error(self._h).decode(errors="replace")

    def __del__(self):
        try:
            self._lib.h264b2_decoder_destroy(self._h)
        except Exception:
            pass
