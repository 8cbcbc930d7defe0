"""Build the CUDA engine (libh264b2.so) in-tree with nvcc for sm_100a.  No JIT, no torch extension:
the product is a plain C-ABI shared library (include/h264_recon_b200.h)."""
import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB = os.path.join(_PKG, "libh264b2.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "residual.cuh", "residual_kernel.cuh", "inter.cuh", "inter_quad.cuh", "inter_tma.cuh", "intra.cuh", "deblock.cuh", "deblock_fast.cuh", "deblock_simd.cuh", "simd16.cuh", "wavefront.cuh"]


HOST_LIB = os.path.join(_PKG, "libh264b2_host.so")
HOST_CLI = os.path.join(_PKG, "h264b2_decode")
HOST_DIR = os.path.join(CSRC, "host")
HOST_SRCS = [os.path.join(HOST_DIR, f) for f in ("H264VideoDecoderB200.cpp", "h264_front.cpp", "h264_slice.cpp", "h264_params.cpp", "h264_multi.cpp")]
HOST_HDRS = [os.path.join(HOST_DIR, f) for f in ("h264_front_internal.h", "h264_decoder.h", "h264_tables.inc")]
CLI_SRC = os.path.join(_ROOT, "tools", "h264b2_decode.cpp")
PARSE_CLI = os.path.join(_PKG, "h264b2_parse")
PARSE_SRC = os.path.join(_ROOT, "tools", "h264b2_parse.cpp")


def build_host(force=False):
    """The C++ host facade (CH264VideoDecoderB200) and its CLI: plain g++ over the C ABI, rpath = $ORIGIN."""
    deps = HOST_SRCS + HOST_HDRS + [CLI_SRC, PARSE_SRC, LIB] + [os.path.join(_ROOT, "include", h) for h in ("H264VideoDecoderB200.h", "h264_recon_b200.h", "h264_front_b200.h", "h264_multi_b200.h")]
    if not force and all(os.path.exists(x) and os.path.getmtime(x) >= max(os.path.getmtime(d) for d in deps) for x in (HOST_LIB, HOST_CLI, PARSE_CLI)):
        return HOST_LIB
    inc = ["-I", os.path.join(_ROOT, "include"), "-I", HOST_DIR]
    common = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-pthread"] + inc
    for cmd in (["g++"] + common + ["-shared", "-o", HOST_LIB] + HOST_SRCS + ["-L", _PKG, "-lh264b2", "-Wl,-rpath,$ORIGIN"],
                ["g++"] + common + ["-o", HOST_CLI, CLI_SRC, "-L", _PKG, "-lh264b2_host", "-lh264b2", "-Wl,-rpath,$ORIGIN"],
                ["g++"] + common + ["-o", PARSE_CLI, PARSE_SRC, "-L", _PKG, "-lh264b2_host", "-lh264b2", "-Wl,-rpath,$ORIGIN"]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building the host facade")
    return HOST_LIB


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 engine cannot be built (there is no CPU fallback)")


def _code_only(text):
    """C/C++ source without comments and with white space collapsed (string and character literals are kept as they are)."""
    out, i, n = [], 0, len(text)
    while i < n:
        c = text[i]
        if c == '"' or c == "'":
            j = i + 1
            while j < n and text[j] != c:
                j += 2 if text[j] == "\\" else 1
            out.append(text[i:j + 1]); i = j + 1
        elif text.startswith("//", i):
            j = text.find("\n", i)
            i = n if j < 0 else j
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            out.append(" "); i = n if j < 0 else j + 2
        else:
            out.append(c); i += 1
    return " ".join("".join(out).split())


def source_hash():
    """SHA-1 over the CODE of the CUDA engine's sources — comments stripped, white space collapsed — because the compiled library is not
    bit-reproducible and a comment must not orphan a profile: what profiles/traffic_r02.json is keyed by."""
    import hashlib
    h = hashlib.sha1()
    for f in sorted([os.path.join(CSRC, x) for x in SOURCES + HEADERS] + [os.path.join(_ROOT, "include", "h264_recon_b200.h")]):
        h.update(os.path.basename(f).encode())
        h.update(_code_only(open(f, encoding="utf-8", errors="replace").read()).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_ROOT, "include", "h264_recon_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_host()
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if verbose else "-warn-spills",
           "-I", os.path.join(_ROOT, "include"), "-I", CSRC] + os.environ.get("H264B2_NVCC_FLAGS", "").split() + [
           "-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES] + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libh264b2.so")
    build_host(force=True)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
