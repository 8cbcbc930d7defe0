"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the Python mirrors of the structs match the C layout byte for byte."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT
from h264_video_decoder_demo_b200 import abi, engine


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "h264_recon_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(h264b2_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libh264b2.so does not export {s}"
    assert set(syms) == set(engine.ABI_SYMBOLS), "engine.py prototypes and the header disagree"
    assert lib.h264b2_abi_version() == abi.ABI_VERSION


def test_host_library_exports_every_front_end_and_decoder_symbol():
    from h264_video_decoder_demo_b200 import frontend
    lib = frontend.lib()
    for hdr, prefix in (("h264_front_b200.h", r"h264b2_front_[a-z0-9_]+"), ("H264VideoDecoderB200.h", r"h264b2_decoder_[a-z0-9_]+")):
        txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", hdr)).read(), flags=re.S)
        syms = sorted(set(re.findall(r"\b(" + prefix + r")\s*\(", txt)))
        assert len(syms) >= 6
        for s in syms:
            assert hasattr(lib, s), f"libh264b2_host.so does not export {s}"


def test_struct_layouts_match_c():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "h264_recon_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(H264B2MbInfo), sizeof(H264B2MbMotion), sizeof(H264B2Weight), sizeof(H264B2PicParams));
  printf("%zu %zu %zu %zu %zu\n", offsetof(H264B2MbInfo, qpy), offsetof(H264B2MbInfo, slice_number), offsetof(H264B2MbInfo, nnz_mask), offsetof(H264B2MbInfo, deblock_idc), offsetof(H264B2MbInfo, coef_mask));
  printf("%zu %zu %zu\n", offsetof(H264B2MbMotion, ref_surf), offsetof(H264B2MbMotion, ref_ident), offsetof(H264B2MbMotion, wt_idx));
  printf("%zu %zu %zu %zu\n", offsetof(H264B2PicParams, dst_surface), offsetof(H264B2PicParams, packed), offsetof(H264B2PicParams, mb_info), offsetof(H264B2PicParams, level_scale8));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split("\n")
    sizes = [int(x) for x in out[0].split()]
    assert sizes == [abi.MB_INFO_DT.itemsize, abi.MB_MOTION_DT.itemsize, abi.WEIGHT_DT.itemsize, C.sizeof(abi.PicParams)]
    f = abi.MB_INFO_DT.fields
    assert [int(x) for x in out[1].split()] == [f["qpy"][1], f["slice_number"][1], f["nnz_mask"][1], f["deblock_idc"][1], f["coef_mask"][1]]
    m = abi.MB_MOTION_DT.fields
    assert [int(x) for x in out[2].split()] == [m["ref_surf"][1], m["ref_ident"][1], m["wt_idx"][1]]
    P = abi.PicParams
    assert [int(x) for x in out[3].split()] == [P.dst_surface.offset, P.packed.offset, P.mb_info.offset, P.level_scale8.offset]


def _pack_model(d):
    """numpy statement of the packed coefficient blob (include/h264_recon_b200.h, k_expand in csrc/engine.cu)"""
    n = d.size
    nc, pad = (n + 15) // 16, (-n) % 16
    ch = np.concatenate([d, np.zeros(pad, np.int16)]).reshape(nc, 16)
    nz = ch != 0
    maps = (nz * (1 << np.arange(16))).sum(axis=1).astype(np.uint16)
    counts = nz.sum(axis=1)
    base = np.concatenate([[0], np.cumsum(counts)])[:-1][::32].astype(np.uint32) if nc else np.zeros(0, np.uint32)
    vals = ch[nz].astype(np.int16)
    body = base.tobytes() + maps.tobytes() + (b"\0\0" if nc & 1 else b"") + vals.tobytes()
    total = 16 + len(body)
    total += (-total) % 16
    hdr = np.array([0x4B503248, nc, vals.size, total], np.uint32).tobytes()
    return (hdr + body).ljust(total, b"\0")


@pytest.mark.parametrize("n", [0, 4, 16, 20, 500, 16 * 32, 16 * 33 + 8, 70000])
@pytest.mark.parametrize("density", [0.0, 0.06, 0.6, 1.0])
def test_packed_coefficient_blob(n, density):
    """h264b2_pack_coefs writes exactly the documented layout, h264b2_unpack_coefs inverts it, the bound always suffices."""
    rng = np.random.default_rng(n + int(density * 100))
    d = (rng.integers(-2000, 2000, n) * (rng.random(n) < density)).astype(np.int16)
    blob = engine.pack_coefs(d)
    assert blob.tobytes() == _pack_model(d)
    assert blob.size <= engine.load_library().h264b2_pack_coefs_bound(n)
    assert np.array_equal(engine.unpack_coefs(blob, n), d)


def test_packed_motion_round_trip():
    """h264b2_pack_motion = XOR chain inside each list + the coefficient packing; unpack inverts it (golden B picture + noise)."""
    from conftest import golden_files
    from h264_video_decoder_demo_b200 import replay
    rp = replay.load_replay(golden_files()[0])
    pic = next(p for p in rp.pictures if p.motion is not None)
    blob = engine.pack_motion(pic.motion)
    assert blob.size < pic.motion.nbytes // 2            # real motion is coarse
    assert engine.unpack_motion(blob, pic.motion.size).tobytes() == pic.motion.tobytes()
    noise = np.frombuffer(np.random.default_rng(5).integers(0, 256, 37 * 152, dtype=np.uint8).tobytes(), dtype=abi.MB_MOTION_DT)
    assert engine.unpack_motion(engine.pack_motion(noise), 37).tobytes() == noise.tobytes()
    d = noise.view(np.uint32).reshape(37, 38).copy()
    for l in range(2):
        d[:, l * 16 + 1:l * 16 + 16] ^= noise.view(np.uint32).reshape(37, 38)[:, l * 16:l * 16 + 15]
    assert engine.pack_motion(noise).tobytes() == _pack_model(d.view(np.int16).reshape(-1))


def test_pack_coefs_reports_a_short_buffer():
    d = np.arange(1, 161, dtype=np.int16)
    raw = np.zeros(64 + 16, np.uint8)
    o = (-raw.ctypes.data) % 16
    with pytest.raises(engine.EngineError):
        engine.pack_coefs(d, raw[o:o + 64])


def test_checksum_host_matches_numpy_definition():
    lib = engine.load_library()
    rng = np.random.default_rng(1)
    for n in (0, 4, 4096, 120 * 68 * 384):
        buf = rng.integers(0, 256, n).astype(np.uint8)
        assert lib.h264b2_checksum_host(buf.ctypes.data, buf.size) == abi.checksum(buf.tobytes())


def test_create_fails_loudly_without_gpu():
    from conftest import _has_gpu
    if _has_gpu():
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.EngineError, match="no CUDA device|no CPU fallback|failed"):
        engine.Engine(0, 1, 8, 6)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "h264_video_decoder_demo_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "recon_oracle" not in txt, f"{fn} references the oracle"
