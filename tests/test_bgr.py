"""Output stage: YUV420P -> BGR24, the reference's integer BT.601 conversion (H264PictureBase.cpp:440-498).

tests/golden/bgr_sums.json holds checksum(reference output frame) -> checksum(the reference's own
convertYuv420pToBgr24 of that frame), written by oracle/ref_harness --bgr-sums.  The CPU test pins the oracle's
restatement to it; the GPU test compares the CUDA kernel with the oracle byte for byte (tolerance 0: integer work)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_files


def _sums():
    with open(os.path.join(GOLDEN_DIR, "bgr_sums.json")) as f:
        return {int(k, 16): int(v, 16) for k, v in json.load(f)["sums"].items()}


def _oracle_frames(path):
    import oracle_py as O
    from h264_video_decoder_demo_b200 import replay
    rp = replay.load_replay(path)
    dpb = O.OracleDPB(rp.width_mbs, rp.height_mbs)
    for pic in rp.pictures:
        dpb.reconstruct(replay.pic_params(rp, pic))
        yield rp, pic, dpb.surfaces[pic.dst_surface]


def test_oracle_bgr24_matches_the_reference_conversion():
    import oracle_py as O
    from h264_video_decoder_demo_b200 import abi
    sums, hits = _sums(), 0
    for path in golden_files():
        for rp, pic, surf in _oracle_frames(path):
            want = sums.get(pic.sum_post)
            if want is None:
                continue
            W, H = rp.width_mbs * 16, rp.height_mbs * 16
            assert abi.checksum(O.convert_bgr24(surf, W, H).tobytes()) == want, f"{os.path.basename(path)} picture {pic.decode_idx}"
            hits += 1
    assert hits >= 3


def test_oracle_bgr24_flip_and_known_pixels():
    import oracle_py as O
    W, H = 4, 2
    i420 = np.array([16, 235, 128, 81, 0, 255, 90, 240] + [128, 90] + [128, 240], dtype=np.uint8)
    a = O.convert_bgr24(i420, W, H).reshape(H, W, 3)
    b = O.convert_bgr24(i420, W, H, flip=True).reshape(H, W, 3)
    assert np.array_equal(a[::-1], b)
    assert a[0, 0].tolist() == [0, 0, 0] and a[0, 1].tolist() == [254, 254, 254]     # (1164*219)/1000 = 254, C truncation
    # U=90, V=240 on the right half: truncation toward zero of a negative intermediate, then clip
    Y, U, V = 1164 * (128 - 16), 90 - 128, 240 - 128
    assert a[0, 2].tolist() == [max(0, min(255, int((Y + 2018 * U) / 1000))), max(0, min(255, int((Y - 813 * V - 391 * U) / 1000))), max(0, min(255, int((Y + 1596 * V) / 1000)))]


@pytest.mark.gpu
def test_gpu_bgr24_bit_exact_vs_oracle_and_reference():
    import oracle_py as O
    from h264_video_decoder_demo_b200 import abi, engine, replay
    sums = _sums()
    path = golden_files()[0]
    rp = replay.load_replay(path)
    W, H = rp.width_mbs * 16, rp.height_mbs * 16
    eng = engine.Engine(0, 1, rp.width_mbs, rp.height_mbs)
    hits = 0
    for pic in rp.pictures:
        eng.submit([0], [replay.pic_params(rp, pic)])
        yuv = eng.read_picture(0, pic.dst_surface)
        got = eng.read_picture_bgr24(0, pic.dst_surface)
        assert np.array_equal(got, O.convert_bgr24(yuv, W, H)), f"picture {pic.decode_idx}"
        if pic.sum_post in sums:
            assert abi.checksum(got.tobytes()) == sums[pic.sum_post]
            hits += 1
    assert hits >= 1
    # bottom-up rows (BMP writer) and a padded row pitch
    flip = eng.read_picture_bgr24(0, rp.pictures[-1].dst_surface, flip_lines=True).reshape(H, W * 3)
    assert np.array_equal(flip[::-1], got.reshape(H, W * 3))
    padded = eng.read_picture_bgr24(0, rp.pictures[-1].dst_surface, width_bytes=W * 3 + 8).reshape(H, W * 3 + 8)
    assert np.array_equal(padded[:, : W * 3], got.reshape(H, W * 3)) and not padded[:, W * 3:].any()
    eng.close()


@pytest.mark.gpu
def test_gpu_bgr24_random_surfaces_small_and_odd_pitch():
    import oracle_py as O
    from h264_video_decoder_demo_b200 import engine
    rng = np.random.default_rng(5)
    for wmb, hmb in ((1, 1), (3, 2), (7, 5)):
        eng = engine.Engine(0, 1, wmb, hmb, surfaces_per_stream=2)
        W, H = wmb * 16, hmb * 16
        surf = rng.integers(0, 256, W * H * 3 // 2, dtype=np.uint8)
        eng.write_picture(0, 1, surf)
        assert np.array_equal(eng.read_picture_bgr24(0, 1), O.convert_bgr24(surf, W, H))
        odd = eng.read_picture_bgr24(0, 1, width_bytes=W * 3 + 1).reshape(H, W * 3 + 1)     # unaligned rows: byte-store path
        assert np.array_equal(odd[:, : W * 3].reshape(-1), O.convert_bgr24(surf, W, H))
        eng.close()
