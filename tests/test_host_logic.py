"""Host-side logic that needs no GPU: replay container round trip, synthetic generator sanity (through the
oracle), stream->rank sharding and the world_size-2 gloo checksum gather used by bench.py --gpus N."""
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT, golden_files
from h264_video_decoder_demo_b200 import replay, sharding


def test_replay_round_trip():
    rp = replay.load_replay(golden_files()[-1])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "x.rp.xz")
        replay.save_replay(rp, path, 2, preset=0)
        rp2 = replay.load_replay(path)
    assert len(rp2.pictures) == 2 and rp2.width_mbs == rp.width_mbs
    for a, b in zip(rp.pictures, rp2.pictures):
        assert a.sum_post == b.sum_post and a.deblock_stop_mb == b.deblock_stop_mb
        assert np.array_equal(a.mb_info, b.mb_info) and np.array_equal(a.coefs, b.coefs)
        assert (a.motion is None) == (b.motion is None)


def test_synthetic_pictures_run_through_the_oracle():
    import oracle_py as O
    import synth
    rng = np.random.default_rng(7)
    for mbaff in (False, True):
        wmb, hmb = 6, 4
        rp = synth.synth_replay(wmb, hmb)
        dpb = O.OracleDPB(wmb, hmb)
        for s in (1, 2):
            dpb.surfaces[s][:] = synth.random_surface(rng, wmb, hmb)
        pic = synth.synth_picture(rng, wmb, hmb, mbaff=mbaff, na_tail=2 if not mbaff else 0)
        dpb.reconstruct(replay.pic_params(rp, pic))
        assert dpb.surfaces[0].any()


def test_shard_plan_is_a_partition():
    for n_units in (1, 5, 64, 77):
        for world in (1, 2, 4, 8):
            plan = [sharding.shard(n_units, r, world) for r in range(world)]
            flat = sorted(u for p in plan for u in p)
            assert flat == list(range(n_units))
            assert max(len(p) for p in plan) - min(len(p) for p in plan) <= 1


def test_gloo_world2_checksum_gather():
    port = socket.socket()
    port.bind(("127.0.0.1", 0))
    p = port.getsockname()[1]
    port.close()
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import torch.distributed as dist\n"
        "from h264_video_decoder_demo_b200 import sharding\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "units = sharding.shard(7, r, w)\n"
        "local = {u: (u * 0x9E3779B97F4A7C15 + 12345) & 0xFFFFFFFFFFFFFFFF for u in units}\n"
        "allsums = sharding.gather_checksums(local, 7)\n"
        "tmax = sharding.max_over_ranks(10.0 + r)\n"
        "assert allsums == [(u * 0x9E3779B97F4A7C15 + 12345) & 0xFFFFFFFFFFFFFFFF for u in range(7)], allsums\n"
        "assert tmax == 11.0, tmax\n"
        "dist.destroy_process_group()\n" % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(p), os.path.join(ROOT, "tests", "_gloo_worker.py")],
                       capture_output=True, text=True, timeout=300, env={**os.environ, "H264B2_TEST_CODE": code})
    assert r.returncode == 0, r.stdout + r.stderr


def _spec_directional(mode, n, x, y, corner, left, top):
    """Directional intra prediction of one sample written from the H.264 equations (8.3.1.2.4-9 / 8.3.2.2.5-10) in p[x, y] notation:
    p(i, -1) = top[i], p(-1, j) = left[j], p(-1, -1) = corner.  n = 4 (Intra_4x4) or 8 (Intra_8x8, on the filtered samples)."""
    def p(i, j):
        if i == -1 and j == -1:
            return corner
        return top[i] if j == -1 else left[j]
    f3 = lambda a, b, c: (a + 2 * b + c + 2) >> 2
    f2 = lambda a, b: (a + b + 1) >> 1
    if mode == 3:
        return f3(p(2 * n - 2, -1), p(2 * n - 1, -1), p(2 * n - 1, -1)) if (x == n - 1 and y == n - 1) else f3(p(x + y, -1), p(x + y + 1, -1), p(x + y + 2, -1))
    if mode == 4:
        if x > y:
            return f3(p(x - y - 2, -1), p(x - y - 1, -1), p(x - y, -1))
        if x < y:
            return f3(p(-1, y - x - 2), p(-1, y - x - 1), p(-1, y - x))
        return f3(p(0, -1), p(-1, -1), p(-1, 0))
    if mode == 5:
        z, k = 2 * x - y, x - (y >> 1)
        if z >= 0 and z % 2 == 0:
            return f2(p(k - 1, -1), p(k, -1))
        if z >= 0:
            return f3(p(k - 2, -1), p(k - 1, -1), p(k, -1))
        if z == -1:
            return f3(p(-1, 0), p(-1, -1), p(0, -1))
        return f3(p(-1, y - 1), p(-1, y - 2), p(-1, y - 3)) if n == 4 else f3(p(-1, y - 2 * x - 1), p(-1, y - 2 * x - 2), p(-1, y - 2 * x - 3))
    if mode == 6:
        z, k = 2 * y - x, y - (x >> 1)
        if z >= 0 and z % 2 == 0:
            return f2(p(-1, k - 1), p(-1, k))
        if z >= 0:
            return f3(p(-1, k - 2), p(-1, k - 1), p(-1, k))
        if z == -1:
            return f3(p(-1, 0), p(-1, -1), p(0, -1))
        return f3(p(x - 1, -1), p(x - 2, -1), p(x - 3, -1)) if n == 4 else f3(p(x - 2 * y - 1, -1), p(x - 2 * y - 2, -1), p(x - 2 * y - 3, -1))
    if mode == 7:
        k = x + (y >> 1)
        return f2(p(k, -1), p(k + 1, -1)) if y % 2 == 0 else f3(p(k, -1), p(k + 1, -1), p(k + 2, -1))
    z, k, lim = x + 2 * y, y + (x >> 1), 2 * n - 3
    if z < lim and z % 2 == 0:
        return f2(p(-1, k), p(-1, k + 1))
    if z < lim:
        return f3(p(-1, k), p(-1, k + 1), p(-1, k + 2))
    if z == lim:
        return (p(-1, n - 2) + 3 * p(-1, n - 1) + 2) >> 2
    return p(-1, n - 1)


@pytest.mark.parametrize("n", [4, 8])
def test_intra_direction_tables_equal_the_prediction_equations(n):
    """The kernel's table-driven directional predictor (modes 3..8) against the equations, for random neighbour samples."""
    import ctypes as C
    from h264_video_decoder_demo_b200 import engine
    tab = np.zeros(6 * n * n, np.uint16)
    assert engine.load_library().h264b2_debug_intra_tables(n, tab.ctypes.data) == 0
    tab = tab.reshape(6, n * n)
    rng = np.random.default_rng(n)
    for _ in range(50):
        corner, left, top = int(rng.integers(0, 256)), [int(v) for v in rng.integers(0, 256, n)], [int(v) for v in rng.integers(0, 256, 2 * n)]
        P = [corner] + left + top                      # the kernel's neighbour array: corner, left[0..n-1], top[0..2n-1]
        for mode in range(3, 9):
            for y in range(n):
                for x in range(n):
                    e = int(tab[mode - 3][y * n + x])
                    a, b, c = P[e & 31], P[(e >> 5) & 31], P[(e >> 10) & 31]
                    got = (a + b + 1) >> 1 if e >> 15 else (a + 2 * b + c + 2) >> 2
                    assert got == _spec_directional(mode, n, x, y, corner, left, top), (n, mode, x, y)
