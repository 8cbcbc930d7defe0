"""Slice short prefixes of the reference-generated replay containers (oracle/_ref/replay/*.bin.xz, made by
oracle/ref_harness from the UNMODIFIED reference) into small committed fixtures under tests/golden/.
Each fixture carries the reference decoder's own pre-/post-deblock picture checksums, so it pins both the
CPU oracle and the CUDA path without /root/reference.   Usage: python tools/make_golden.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h264_video_decoder_demo_b200 import replay  # noqa: E402

# (stream, pictures): tff needs 7 so that the one picture with field macroblocks (decode index 6) is covered
PLAN = [("HeavyHand_1080p.B_frames.cabac", 4), ("HeavyHand_1080p.B_frames_4.no_cabac.no_tff", 3),
        ("HeavyHand_1080p.B_frames_cabac_tff", 7), ("HeavyHand_1080p.no_B_frames.cabac.no_tff", 3), ("gop121.naluCnt453", 3)]


def make_bitstream_prefixes(out_dir):
    """Annex-B prefixes of the bundled streams holding the first N pictures (same N as the container fixtures) plus the
    first slice NAL of picture N+1, so that picture N completes exactly as it does in the full stream (deblocked, marked)."""
    for name, n in PLAN:
        d = open(os.path.join(ROOT, "oracle", "_ref", "streams", name + ".h264"), "rb").read()
        starts, i = [], 0
        while True:
            j = d.find(b"\x00\x00\x01", i)
            if j < 0:
                break
            starts.append(j + 3)
            i = j + 3
        pics, cut = 0, None
        for k, p in enumerate(starts):
            if (d[p] & 31) in (1, 5) and (d[p + 1] & 0x80):        # a slice NAL with first_mb_in_slice == 0: a new picture
                pics += 1
                if pics == n + 1:
                    cut = starts[k + 1] - 3 if k + 1 < len(starts) else len(d)
                    while cut > p and d[cut - 1] == 0:
                        cut -= 1
                    break
        dst = os.path.join(out_dir, f"{name}.first{n}.h264")
        with open(dst, "wb") as f:
            f.write(d[:cut])
        print(dst, cut)


SYNTH = [("synth_cavlc_ip", dict(seed=11, n_pics=5)),
         ("synth_t8x8_wp_poc0", dict(seed=12, t8x8=True, weighted=True, n_refs=4, poc_type=0, n_pics=6)),
         ("synth_slices_11x9", dict(seed=13, wmb=11, hmb=9, max_slices=5, n_pics=4)),
         ("synth_wp_5x4", dict(seed=14, weighted=True, wmb=5, hmb=4, n_pics=8, n_refs=4)),
         ("synth_b_direct", dict(seed=15, bframes=True, n_pics=7)),
         ("synth_b_implicit", dict(seed=16, bframes=True, bipred_idc=2, n_pics=7, n_refs=4)),
         ("synth_b_explicit_t8x8", dict(seed=17, bframes=True, bipred_idc=1, weighted=True, n_pics=7, t8x8=True)),
         ("synth_b_explicit_6x5", dict(seed=18, bframes=True, bipred_idc=1, wmb=6, hmb=5, n_pics=9, max_slices=1)),
         # adaptive reference marking, long-term reference pictures, list modification with long-term picture numbers.
         # What the UNMODIFIED reference decodes: MMCO 1, 2, 3, 5, 6, IDR long_term_reference_flag, modification_of_pic_nums_idc 0/1/2.
         # MMCO 4 (max_long_term_frame_idx) makes the reference crash or lose its reference lists (tests/test_front_end.py records it).
         ("synth_mmco1_mod", dict(seed=301, mmco=True, mmco_set=(1,), mmco_idr_lt=False, n_pics=8, n_refs=4, max_slices=2)),
         ("synth_mmco_lt_idr", dict(seed=315, mmco=True, mmco_set=(1, 2), n_pics=8, n_refs=4, max_slices=2)),
         ("synth_mmco_lt_36", dict(seed=331, mmco=True, mmco_set=(1, 2, 3, 6), n_pics=9, n_refs=4, max_slices=2)),
         ("synth_mmco_lt_poc0_wp", dict(seed=300, mmco=True, mmco_set=(1, 2, 3, 6), poc_type=0, weighted=True, t8x8=True, n_pics=8, n_refs=4, max_slices=2)),
         ("synth_mmco5_poc0", dict(seed=301, mmco=True, mmco_set=(1,), mmco_idr_lt=False, mmco5=True, poc_type=0, n_pics=9, n_refs=4, max_slices=2))]


def make_synthetic(out_dir):
    """Random-syntax CAVLC streams (tests/h264_writer.py) decoded by the UNMODIFIED reference (oracle/_ref/ref_harness): the stream and
    the reference's structure-of-arrays + picture checksums become fixtures.  They reach what the bundled streams never do: I_PCM,
    P sub-partitions 8x4/4x8/4x4, P_8x8ref0, every Intra16x16 mode, several slices per picture with idc 0/1/2 and offsets,
    per-slice explicit weights, CAVLC with the 8x8 transform; B pictures (reordered output, B pyramid) with every 16x16/16x8/8x16 list
    combination, B_8x8 incl. direct sub-macroblocks, spatial direct, implicit and EXPLICIT bi-prediction weights (Q8 live)."""
    import lzma
    import subprocess
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import h264_writer
    only = os.environ.get("GOLDEN_ONLY")            # e.g. GOLDEN_ONLY=synth_mmco: regenerate just those fixtures
    for name, cfg in SYNTH:
        if only and not name.startswith(only):
            continue
        data = h264_writer.Stream(**cfg).build()
        n = cfg["n_pics"]
        with open(os.path.join(out_dir, f"{name}.first{n}.h264"), "wb") as f:
            f.write(data)
        with tempfile.NamedTemporaryFile(suffix=".bin") as t:
            r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), os.path.join(out_dir, f"{name}.first{n}.h264"), "--replay", t.name, "--quiet"],
                               capture_output=True, text=True)
            bad = [l for l in r.stdout.split("\n") if ("failed" in l or "Error" in l) and "open: Error" not in l]
            assert not bad, (name, bad[:3])
            blob = open(t.name, "rb").read()
        dst = os.path.join(out_dir, f"{name}.first{n}.rp.xz")
        with lzma.open(dst, "wb", preset=9) as f:
            f.write(blob)
        print(dst, len(data), os.path.getsize(dst))


def make_bgr_sums(out_dir):
    """checksum(reference YUV output frame) -> checksum(reference convertYuv420pToBgr24 of it), first 3 output frames of 3 streams."""
    import json
    import subprocess
    import tempfile
    ref = os.path.join(ROOT, "oracle", "_ref")
    sums = {}
    for name in ("HeavyHand_1080p.B_frames.cabac", "HeavyHand_1080p.B_frames_cabac_tff", "gop121.naluCnt453"):
        with tempfile.NamedTemporaryFile("r", suffix=".txt") as t:
            subprocess.run([os.path.join(ref, "ref_harness"), os.path.join(ref, "streams", name + ".h264"), "--quiet", "--max-frames", "3", "--bgr-sums", t.name],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for line in open(t.name):
                _, _, yuv, bgr = line.split()
                sums[yuv] = bgr
    note = ("checksum(reference YUV frame) -> checksum(reference CH264PictureBase::convertYuv420pToBgr24 of that frame, stride W*3); "
            "generated by oracle/ref_harness --bgr-sums on the first 3 output frames of 3 bundled streams")
    with open(os.path.join(out_dir, "bgr_sums.json"), "w") as f:
        json.dump({"note": note, "sums": sums}, f, indent=1)


if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    make_bgr_sums(out_dir)
    make_bitstream_prefixes(out_dir)
    make_synthetic(out_dir)
    for name, n in PLAN:
        rp = replay.load_replay(os.path.join(replay.default_replay_dir(), name + ".bin.xz"), n)
        dst = os.path.join(out_dir, f"{name}.first{n}.rp.xz")
        replay.save_replay(rp, dst, n, preset=9)
        print(dst, os.path.getsize(dst))
