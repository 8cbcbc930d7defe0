"""Pin the C restatement (recon_oracle.c) against the reference: run it over every picture of a
replay file and compare pre-/post-deblock checksums with the ones ref_harness recorded from the
unmodified reference.  TEST INFRASTRUCTURE.  Usage: python oracle/pin_check.py [replay files...]"""
import glob
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def check(path, max_pictures=None):
    import oracle_py as O
    from h264_video_decoder_demo_b200 import replay
    rp = replay.load_replay(path, max_pictures)
    dpb = O.OracleDPB(rp.width_mbs, rp.height_mbs)
    bad = []
    t0 = time.time()
    sums = {}
    for pic in rp.pictures:
        p = replay.pic_params(rp, pic)
        dpb.reconstruct(p, O.STAGE_RECON)
        pre = dpb.checksum(pic.dst_surface)
        dpb.reconstruct(p, O.STAGE_DEBLOCK)
        post = dpb.checksum(pic.dst_surface)
        sums[pic.decode_idx] = post
        if pre != pic.sum_pre or post != pic.sum_post:
            bad.append((pic.decode_idx, pre == pic.sum_pre, post == pic.sum_post))
    out_ok = all(sums.get(i) == s for i, s in zip(rp.out_order, rp.out_sums) if i in sums)
    return os.path.basename(path), len(rp.pictures), bad, out_ok, time.time() - t0


if __name__ == "__main__":
    files = sys.argv[1:] or sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "replay", "*.bin.xz")))
    with ProcessPoolExecutor(max_workers=min(8, len(files))) as ex:
        for name, n, bad, out_ok, dt in ex.map(check, files):
            print(f"{name}: {n} pictures, mismatches={bad[:8]}{'...' if len(bad) > 8 else ''} ({len(bad)}), output-order sums ok={out_ok}, {dt:.1f}s")
