"""GPU parity tests proper: the CUDA path, called through the C ABI (libh264b2.so), must be BIT-EXACT
(integer/byte work: tolerance 0) against
  (1) the unmodified reference decoder's recorded picture checksums (golden fixtures + full replays), and
  (2) the CPU oracle on seeded synthetic pictures that reach cases the bundled streams never do.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import full_files, golden_files
from h264_video_decoder_demo_b200 import abi, engine, replay

pytestmark = pytest.mark.gpu


def _run_stream(eng, rs, sid=0, check_pre=False):
    bad = []
    for i, pic in enumerate(rs.rp.pictures):
        p = rs.params[i]
        if check_pre:
            saved = p.deblock_enable
            p.deblock_enable = 0
            eng.submit_device([sid], [p])
            if eng.checksum(sid, pic.dst_surface) != pic.sum_pre:
                bad.append(("pre", i))
            p.deblock_enable = saved
        eng.submit_device([sid], [p])
        if eng.checksum(sid, pic.dst_surface) != pic.sum_post:
            bad.append(("post", i))
    return bad


@pytest.mark.parametrize("path", golden_files("all"), ids=os.path.basename)
def test_golden_fixture_bit_exact(path):
    rp = replay.load_replay(path)
    eng = engine.Engine(0, 1, rp.width_mbs, rp.height_mbs)
    rs = engine.ResidentStream(eng, rp)
    assert _run_stream(eng, rs, check_pre=True) == []
    # output-order checksums (what the callback would see) for the frames this prefix completes
    eng.close()


@pytest.mark.parametrize("path", full_files(), ids=os.path.basename)
def test_full_stream_bit_exact(path):
    """Every picture of the bundled stream, pre-parsed by the reference's own parser, reconstructed on the GPU."""
    rp = replay.load_replay(path)
    eng = engine.Engine(0, 1, rp.width_mbs, rp.height_mbs)
    rs = engine.ResidentStream(eng, rp)
    assert _run_stream(eng, rs) == []
    eng.close()


@pytest.mark.parametrize("packed", [False, True], ids=["plain_arrays", "packed_arrays"])
def test_all_streams_in_one_batch_and_host_submit(packed):
    """Five different streams share every launch (mixed I/P/B, MBAFF and non-MBAFF pictures in one batch);
    arrays are passed as HOST pointers through h264b2_submit — with dense coefficient arrays or with the packed
    transport (h264b2_pack_coefs -> k_expand); frames come back through the async read path."""
    files = golden_files()
    rps = [replay.load_replay(f) for f in files]
    n = len(rps)
    eng = engine.Engine(0, n, rps[0].width_mbs, rps[0].height_mbs)
    depth = max(len(r.pictures) for r in rps)
    host = eng.pinned_array(n * eng.frame_bytes).reshape(n, eng.frame_bytes)
    for i in range(depth):
        sids = [s for s in range(n) if i < len(rps[s].pictures)]
        blobs = [engine.pack_coefs(rps[s].pictures[i].coefs) if packed else None for s in sids]
        mblobs = [engine.pack_motion(rps[s].pictures[i].motion) if packed and rps[s].pictures[i].motion is not None else None for s in sids]
        params = [replay.pic_params(rps[s], rps[s].pictures[i], packed_blob=b, packed_motion=m) for s, b, m in zip(sids, blobs, mblobs)]
        eng.submit(sids, params)
        eng.read_pictures_async(sids, [rps[s].pictures[i].dst_surface for s in sids], [host[s].ctypes.data for s in sids])
        eng.sync()
        for s in sids:
            assert abi.checksum(host[s].tobytes()) == rps[s].pictures[i].sum_post, f"stream {s} picture {i}"
        sums = eng.checksums(sids, [rps[s].pictures[i].dst_surface for s in sids])
        assert sums == [rps[s].pictures[i].sum_post for s in sids]
    eng.close()


def _describe(got, ref, pic, wmb, hmb):
    W, H = wmb * 16, hmb * 16
    d = np.nonzero(got != ref)[0]
    names = {0: "NA", 1: "I4x4", 2: "I8x8", 3: "I16x16", 4: "IPCM", 5: "INTER"}
    out = [f"{d.size} bytes differ"]
    seen = set()
    for off in d:
        off = int(off)
        if off < W * H:
            comp, x, y, mbs, w = "Y", off % W, off // W, 16, W
        else:
            o2 = off - W * H
            comp = "Cb" if o2 < W * H // 4 else "Cr"
            o2 %= W * H // 4
            x, y, mbs, w = o2 % (W // 2), o2 // (W // 2), 8, W // 2
        mbx, mby = x // mbs, y // mbs
        if pic.mbaff:
            pr = (mby // 2) * wmb + mbx
            addrs = [2 * pr, 2 * pr + 1]
        else:
            addrs = [mby * wmb + mbx]
        key = (comp, tuple(addrs))
        if key in seen:
            continue
        seen.add(key)
        info = "; ".join(f"a={a} {names[int(pic.mb_info['mb_class'][a])]} flags={int(pic.mb_info['flags'][a]):#x} slice={int(pic.mb_info['slice_number'][a])} cm={int(pic.mb_info['coef_mask'][a]):#x}" for a in addrs)
        out.append(f"{comp}({x},{y}) gpu={got[off]} oracle={ref[off]} MB({mbx},{mby}) [{info}]")
        if len(seen) >= 6:
            break
    return " | ".join(out)


def _oracle_vs_gpu(rng, wmb, hmb, pics_kwargs, smooth):
    import oracle_py as O
    import synth
    rp = synth.synth_replay(wmb, hmb)
    eng = engine.Engine(0, 1, wmb, hmb, surfaces_per_stream=4)
    dpb = O.OracleDPB(wmb, hmb, n_surfaces=4)
    for s in (1, 2, 3):
        surf = synth.random_surface(rng, wmb, hmb, smooth=smooth)
        dpb.surfaces[s][:] = surf
        eng.write_picture(0, s, surf)
    stale = synth.random_surface(rng, wmb, hmb, smooth=smooth)     # Q15: what an unpredicted block keeps
    for kw in pics_kwargs:
        pic = synth.synth_picture(rng, wmb, hmb, **kw)
        dpb.surfaces[0][:] = stale
        eng.write_picture(0, 0, stale)
        p = replay.pic_params(rp, pic)
        for stages, dbk in ((O.STAGE_RECON, 0), (O.STAGE_RECON | O.STAGE_DEBLOCK, 1)):
            dpb.surfaces[0][:] = stale
            eng.write_picture(0, 0, stale)
            p.deblock_enable = dbk and pic.deblock_enable
            dpb.reconstruct(p, stages)
            # the deblocked pass travels with packed coefficients (incl. I_PCM samples and custom scaling lists)
            q = p
            if dbk:
                blob = engine.pack_coefs(pic.coefs)
                mblob = engine.pack_motion(pic.motion) if pic.motion is not None else None
                q = replay.pic_params(rp, pic, packed_blob=blob, packed_motion=mblob)
                q.deblock_enable = p.deblock_enable
            eng.submit([0], [q])
            got = eng.read_picture(0, 0)
            if not np.array_equal(got, dpb.surfaces[0]):
                raise AssertionError(f"{kw} deblock={dbk}: " + _describe(got, dpb.surfaces[0], pic, wmb, hmb))
    eng.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_synthetic_progressive_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    cases = [dict(inter_frac=0.0, amp=40), dict(inter_frac=1.0, pcm_frac=0.0, amp=30), dict(inter_frac=0.6, amp=8, n_slices=4),
             dict(inter_frac=0.5, custom_scaling=True, amp=20), dict(inter_frac=0.5, na_tail=5, amp=10), dict(inter_frac=0.5, cip=True, amp=5),
             dict(inter_frac=0.7, amp=2000, n_slices=1)]
    _oracle_vs_gpu(rng, 9, 7, cases, smooth=bool(seed & 1))


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_synthetic_mbaff_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    cases = [dict(mbaff=True, inter_frac=0.0, amp=30), dict(mbaff=True, inter_frac=1.0, pcm_frac=0.0, amp=10),
             dict(mbaff=True, inter_frac=0.5, amp=4, n_slices=3), dict(mbaff=True, inter_frac=0.5, cip=True, custom_scaling=True, amp=12)]
    _oracle_vs_gpu(rng, 7, 6, cases, smooth=bool(seed & 1))


def test_edge_sizes_vs_oracle():
    rng = np.random.default_rng(11)
    _oracle_vs_gpu(rng, 1, 1, [dict(inter_frac=0.5, n_slices=1), dict(inter_frac=0.0, n_slices=1)], smooth=True)
    _oracle_vs_gpu(rng, 1, 2, [dict(mbaff=True, inter_frac=0.5, n_slices=1)], smooth=True)
    _oracle_vs_gpu(rng, 40, 1, [dict(inter_frac=0.5, n_slices=2)], smooth=True)
    _oracle_vs_gpu(rng, 2, 34, [dict(inter_frac=0.5, n_slices=2), dict(mbaff=True, inter_frac=0.3)], smooth=False)


def test_full_size_idempotence_and_replica_independence():
    """1080p, size-independent properties: reconstructing the same pictures again gives the same bytes
    (no state leaks between submits), and replicas of one stream in one batch all agree (no cross-stream
    interference in the shared wavefront launches)."""
    rp = replay.load_replay(golden_files()[0])
    S = 6
    eng = engine.Engine(0, S, rp.width_mbs, rp.height_mbs)
    rs = engine.ResidentStream(eng, rp)
    for rep in range(2):
        for i, pic in enumerate(rp.pictures):
            eng.submit_device(list(range(S)), [rs.params[i]] * S)
            sums = eng.checksums(list(range(S)), [pic.dst_surface] * S)
            assert sums == [pic.sum_post] * S, f"rep {rep} picture {i}: {sums}"
    eng.close()


def test_error_behaviour():
    rp = replay.load_replay(golden_files()[-1], 1)
    eng = engine.Engine(0, 2, rp.width_mbs, rp.height_mbs)
    p = replay.pic_params(rp, rp.pictures[0])
    with pytest.raises(engine.EngineError):
        eng.submit([0, 0], [p, p])                   # one stream twice in a batch
    with pytest.raises(engine.EngineError):
        eng.submit([5], [p])                         # stream out of range
    p.dst_surface = 99
    with pytest.raises(engine.EngineError):
        eng.submit([0], [p])
    p = replay.pic_params(rp, rp.pictures[0])
    p.width_mbs = 8
    with pytest.raises(engine.EngineError):
        eng.submit([0], [p])
    p = replay.pic_params(rp, rp.pictures[0])
    p.packed = 1                                     # claims a packed blob, points at dense levels
    with pytest.raises(engine.EngineError):
        eng.submit([0], [p])
    p.packed = 8
    with pytest.raises(engine.EngineError):
        eng.submit([0], [p])
    eng.close()


def test_closed_gops_of_one_stream_decode_in_parallel():
    """closed GOPs shard like independent streams.  The two GOPs of a HeavyHand stream (IDRs at pictures
    0 and 46) are reconstructed CONCURRENTLY on two DPBs of one context; every picture still equals the reference's
    full-stream decode (the GOP-0 tail picture is deblocked there, and so it is here)."""
    from h264_video_decoder_demo_b200 import sharding
    files = [f for f in full_files() if "B_frames.cabac" in f]
    if not files:
        pytest.skip("full replays not built")
    rp = replay.load_replay(files[0])
    starts = sharding.closed_gop_starts(rp)
    assert starts == [0, 46]
    bounds = starts + [len(rp.pictures)]
    eng = engine.Engine(0, len(starts), rp.width_mbs, rp.height_mbs)
    rs = engine.ResidentStream(eng, rp)
    depth = max(bounds[g + 1] - bounds[g] for g in range(len(starts)))
    for k in range(depth):
        sids = [g for g in range(len(starts)) if bounds[g] + k < bounds[g + 1]]
        idx = [bounds[g] + k for g in sids]
        eng.submit_device(sids, [rs.params[i] for i in idx])
        assert eng.checksums(sids, [rp.pictures[i].dst_surface for i in idx]) == [rp.pictures[i].sum_post for i in idx], f"step {k}"
    eng.close()
