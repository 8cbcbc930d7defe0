"""GPU parity driver + diagnosis (TEST INFRASTRUCTURE; uses the CPU oracle as the checker).

python tests/parity_tool.py [--max N] [--stage pre|post] replay files...
Runs every picture of each replay through the CUDA engine (one stream) and compares the picture checksum with
the one recorded from the unmodified reference.  On the first mismatch, re-runs that picture through the
oracle (seeded with the GPU's reference surfaces) and reports which macroblocks differ.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from h264_video_decoder_demo_b200 import abi, engine, replay  # noqa: E402


def diagnose(eng, rp, pic, stage_deblock):
    import oracle_py as O
    dpb = O.OracleDPB(rp.width_mbs, rp.height_mbs)
    gpu = eng.read_picture(0, pic.dst_surface)
    for s in range(17):
        if s != pic.dst_surface:
            dpb.surfaces[s][:] = eng.read_picture(0, s)
    p = replay.pic_params(rp, pic)
    if not stage_deblock:
        p.deblock_enable = 0
    dpb.reconstruct(p, O.STAGE_RECON | (O.STAGE_DEBLOCK if stage_deblock else 0))
    ref = dpb.surfaces[pic.dst_surface]
    W, H = rp.width_mbs * 16, rp.height_mbs * 16
    names = {0: "NA", 1: "I4x4", 2: "I8x8", 3: "I16x16", 4: "IPCM", 5: "INTER"}
    planes = [("Y", 0, W, H, 16), ("Cb", W * H, W // 2, H // 2, 8), ("Cr", W * H + W * H // 4, W // 2, H // 2, 8)]
    for name, off, w, h, mbs in planes:
        g = gpu[off:off + w * h].reshape(h, w).astype(int)
        r = ref[off:off + w * h].reshape(h, w).astype(int)
        d = (g != r)
        print(f"  plane {name}: {int(d.sum())} differing samples, max |diff| {int(np.abs(g - r).max())}")
        if d.any():
            ys, xs = np.nonzero(d)
            mbx, mby = xs // mbs, ys // mbs
            seen = []
            for x, y in zip(mbx, mby):
                if (x, y) not in seen:
                    seen.append((x, y))
                if len(seen) >= 12:
                    break
            for (x, y) in seen:
                if pic.mbaff:
                    addrs = [2 * ((y // 2) * rp.width_mbs + x), 2 * ((y // 2) * rp.width_mbs + x) + 1]
                else:
                    addrs = [y * rp.width_mbs + x]
                desc = ", ".join(f"a={a} {names[int(pic.mb_info['mb_class'][a])]} flags={int(pic.mb_info['flags'][a]):#x} cm={int(pic.mb_info['coef_mask'][a]):#x}" for a in addrs)
                sel = d[y * mbs:(y + 1) * mbs, x * mbs:(x + 1) * mbs]
                yy, xx = np.nonzero(sel)
                print(f"    MB({x},{y}) {desc}: {int(sel.sum())} px, first at ({xx[0]},{yy[0]}) gpu={g[y*mbs+yy[0], x*mbs+xx[0]]} ref={r[y*mbs+yy[0], x*mbs+xx[0]]}")


def run(path, max_pics, stage):
    t0 = time.time()
    rp = replay.load_replay(path, max_pics)
    eng = engine.Engine(0, 1, rp.width_mbs, rp.height_mbs)
    rs = engine.ResidentStream(eng, rp)
    print(f"{os.path.basename(path)}: {len(rp.pictures)} pictures loaded in {time.time() - t0:.1f}s")
    bad = 0
    t0 = time.time()
    for i, pic in enumerate(rp.pictures):
        p = rs.params[i]
        if stage == "pre":
            saved = p.deblock_enable
            p.deblock_enable = 0
            eng.submit_device([0], [p])
            s = eng.checksum(0, pic.dst_surface)
            p.deblock_enable = saved
            if s != pic.sum_pre:
                print(f" picture {i} (type {pic.slice_type} mbaff {pic.mbaff}): PRE-deblock mismatch")
                diagnose(eng, rp, pic, False)
                bad += 1
                break
        eng.submit_device([0], [p])
        s = eng.checksum(0, pic.dst_surface)
        if s != pic.sum_post:
            print(f" picture {i} (type {pic.slice_type} mbaff {pic.mbaff} deblock {pic.deblock_enable}): POST-deblock mismatch")
            diagnose(eng, rp, pic, True)
            bad += 1
            break
    print(f"  {'OK' if not bad else 'FAIL'}: {i + 1} pictures checked in {time.time() - t0:.1f}s")
    rs.free()
    eng.close()
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="*")
    ap.add_argument("--max", type=int, default=None)
    ap.add_argument("--stage", default="pre")
    a = ap.parse_args()
    files = a.files or sorted(os.path.join(replay.default_replay_dir(), f) for f in os.listdir(replay.default_replay_dir()) if f.endswith(".bin.xz"))
    rc = 0
    for f in files:
        rc |= run(f, a.max, a.stage)
    sys.exit(1 if rc else 0)
