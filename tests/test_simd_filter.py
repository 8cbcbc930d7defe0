"""CPU check of the packed two-line deblocking filter the CUDA kernel k_deblock3 runs (csrc/simd16.cuh compiles for the host too):
random sample patches, thresholds and boundary strengths through the packed filter and the 4x4 block re-formatting must equal the
scalar filter equations of the reference (H264PictureDeblockingFilterProcess.cpp:1314-1522) sample for sample."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_deblock_filter_equals_the_scalar_equations(tmp_path):
    exe = str(tmp_path / "simd_filter_check")
    src = os.path.join(ROOT, "tests", "simd_filter_check.cpp")
    inc = os.path.join(ROOT, "h264_video_decoder_demo_b200", "csrc")
    subprocess.run(["g++", "-O2", "-x", "c++", "-I", inc, "-o", exe, src], check=True)
    r = subprocess.run([exe, "1500000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    fields = r.stdout.split()
    assert int(fields[fields.index("bad") + 1]) == 0
    assert int(fields[fields.index("changed") + 1]) > 100000      # the patches really exercise the filters
