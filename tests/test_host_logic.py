"""Host-side logic that needs no GPU: replay container round trip, synthetic generator sanity (through the
oracle), stream->rank sharding and the world_size-2 gloo checksum gather used by bench.py --gpus N."""
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np

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
