"""include/H264VideoDecoder.h — the same-name shim of the reference's decoder class (H264VideoDecoder.h:22-43): a client written like the
reference's main.cpp compiles against it and, on a GPU, produces the reference's frames; where /root/reference is mounted, the
reference's OWN main.cpp is compiled against the shim as it stands."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT

PKG = os.path.join(ROOT, "h264_video_decoder_demo_b200")


def _build_client(tmp_path):
    exe = str(tmp_path / "shim_main_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "shim_main_check.cpp"),
                    "-L", PKG, "-lh264b2_host", "-lh264b2", f"-Wl,-rpath,{PKG}"], check=True)
    return exe


def test_client_written_against_the_reference_interface_builds_against_the_shim(tmp_path):
    assert os.path.exists(_build_client(tmp_path))


def test_the_reference_main_cpp_compiles_against_the_shim(tmp_path):
    ref = os.environ.get("H264B2_REFERENCE", "/root/reference")
    src = os.path.join(ref, "h264_video_decoder_demo", "main.cpp")
    if not os.path.exists(src):
        pytest.skip("reference sources not mounted")
    # a scratch copy: the quoted include must find OUR H264VideoDecoder.h, not the one next to the reference's main.cpp
    shutil.copy(src, tmp_path / "main.cpp")
    shutil.copy(os.path.join(ref, "h264_video_decoder_demo", "version.h"), tmp_path / "version.h")
    subprocess.run(["g++", "-std=c++17", "-w", "-c", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "main.o"), str(tmp_path / "main.cpp")], check=True)
    subprocess.run(["g++", "-o", str(tmp_path / "demo"), str(tmp_path / "main.o"), "-L", PKG, "-lh264b2_host", "-lh264b2", f"-Wl,-rpath,{PKG}"], check=True)


@pytest.mark.gpu
def test_shim_client_decodes_to_the_reference_frames(tmp_path):
    import oracle_py as O
    from h264_video_decoder_demo_b200 import replay
    exe = _build_client(tmp_path)
    stem = os.path.join(GOLDEN_DIR, "synth_b_direct.first7")
    out = tmp_path / "bmp"
    out.mkdir()
    r = subprocess.run([exe, stem + ".h264", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rp = replay.load_replay(stem + ".rp.xz")
    assert f"frames={len(rp.out_order)} end_seen=1" in r.stdout
    assert "profile=100" in r.stdout and "cabac=0" in r.stdout and "fps=25.000" in r.stdout
    # frame 0 of the output order, as the BMP holds it (bottom-up BGR rows), equals the oracle's reconstruction of that picture
    dpb = O.OracleDPB(rp.width_mbs, rp.height_mbs)
    surf = {}
    for pic in rp.pictures:
        dpb.reconstruct(replay.pic_params(rp, pic))
        surf[pic.decode_idx] = dpb.surfaces[pic.dst_surface].copy()
    W, H = rp.width_mbs * 16, rp.height_mbs * 16
    bmp = np.fromfile(str(out / f"out_{W}x{H}.0.bmp"), dtype=np.uint8)
    pitch = (W * 3 + 3) & ~3
    rows = bmp[54:].reshape(H, pitch)[::-1, : W * 3]
    assert np.array_equal(rows.reshape(-1), O.convert_bgr24(surf[rp.out_order[0]], W, H))
