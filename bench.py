#!/usr/bin/env python
"""bench.py — 1080p pictures/s of the B200 H.264 picture-reconstruction engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--workload NAME] [--impl reference]

A "step" is one pass of the hot path over one whole bundled stream for S concurrent replicas on every GPU:
76 batched submits (one per picture in decoding order; pictures of one stream are serial), each
reconstructing S pictures.  Inputs are the per-picture structure-of-arrays that the host entropy stage
emits (pre-parsed by the unmodified reference's own parser through oracle/ref_harness), made resident in
HBM before the timed region — one distinct device copy per replica, 16 GB at S=64, far larger than L2.

  value     pictures/s, device-timed (CUDA events on the launch stream), inputs resident in HBM
  e2e       pictures/s through h264b2_submit() with HOST (pinned) buffers: H2D of every picture's SoA and D2H
            of every reconstructed picture (what the reference's output callback receives) inside the region
  roofline  dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference decoder (oracle/_ref/ref_harness) on one host core, bounded sample

--impl reference times the reference's own CPU decoder on all usable host cores (one process per core).
Multi-GPU (torchrun, one rank per GPU): streams shard with no data-path collective; weak scaling.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "1080p frames/sec decoded (picture reconstruction), YUV bit-exact vs reference"
UNIT = "frames/s"
WORKLOADS = {
    "B_frames.cabac": "HeavyHand_1080p.B_frames.cabac",
    "B_frames_4.no_cabac": "HeavyHand_1080p.B_frames_4.no_cabac.no_tff",
    "tff": "HeavyHand_1080p.B_frames_cabac_tff",
    "no_B_frames.cabac": "HeavyHand_1080p.no_B_frames.cabac.no_tff",
    "gop121": "gop121.naluCnt453",
}
MIXED = ["B_frames.cabac", "B_frames_4.no_cabac", "tff", "no_B_frames.cabac", "gop121"]     # BASELINE config 5: all bundled variants, round-robin
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def find_replay(stem):
    full = os.path.join(REF_DIR, "replay", stem + ".bin.xz")
    if os.path.exists(full):
        return full, True
    gd = os.path.join(ROOT, "tests", "golden")
    for f in sorted(os.listdir(gd)):
        if f.startswith(stem + ".first"):
            return os.path.join(gd, f), False
    raise FileNotFoundError(f"no replay container for {stem}")


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[2:6]):
                if v == "Active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ reference CPU decoder
def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_once(stream_stem, max_frames, nproc):
    """nproc single-threaded reference decoders in parallel, each decoding the first pictures of the stream until
    max_frames frames have reached the output callback.  Returns (pictures decoded in total, wall seconds)."""
    exe = os.path.join(REF_DIR, "ref_harness")
    src = os.path.join(REF_DIR, "streams", stream_stem + ".h264")
    if not (os.path.exists(exe) and os.path.exists(src)):
        return None
    t0 = time.time()
    procs = [subprocess.Popen([exe, src, "--max-frames", str(max_frames), "--quiet"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
             for _ in range(nproc)]
    pics = 0
    for p in procs:
        _, err = p.communicate()
        m = re.search(r"RESULT .*frames=(\d+) pics=(\d+) secs=([\d.]+)", err or "")
        if m:
            pics += int(m.group(2))
    return pics, time.time() - t0


def reference_arm(args, stem):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(os.path.join(REF_DIR, "ref_harness")):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness not built (needs /root/reference at build time)"}))
        return
    cores = usable_cores()
    try:
        avail_gb = int(re.search(r"MemAvailable:\s+(\d+)", open("/proc/meminfo").read()).group(1)) / 1048576
    except Exception:
        avail_gb = 16
    nproc = max(1, min(cores, int(avail_gb * 0.5 / 1.5), 64))       # ~1 GB RSS per decoder on a short sample
    total_steps = args.steps + args.warmup
    frames = max(3, min(10, int(150.0 / max(total_steps, 1) * 1.2)))   # keep the whole run to a few minutes
    for _ in range(args.warmup):
        run_reference_once(stem, frames, nproc)
    pics, secs = 0, 0.0
    for _ in range(args.steps):
        p, s = run_reference_once(stem, frames, nproc)
        pics += p; secs += s
    v = pics / secs
    sample = f"{nproc} parallel single-threaded reference decoders x first {frames} output frames of {stem}.h264 per step (process start-up included)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1000 * secs / max(args.steps, 1), 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "bundled reference bitstream", "config": {"workload": stem + ".h264 (reference CPU decoder, all usable host cores)"},
        "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": nproc, "kind": "reference", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------ algorithmic bytes
def algorithmic_bytes(rp):
    """Per picture: MC = 24 B x (nLists + 1) per inter 4x4 luma block (16 Y + 4 Cb + 4 Cr read once per list,
    written once); deblock = every sample of a deblocked picture read + written once; intra = samples written."""
    import numpy as np
    from h264_video_decoder_demo_b200 import abi
    out = []
    for p in rp.pictures:
        mc = 0
        if p.motion is not None:
            inter = p.mb_info["mb_class"] == abi.MB_INTER
            lists = (p.motion["ref_surf"][inter] >= 0).sum(axis=1)      # [n_inter][4 quadrants] -> lists per quadrant
            mc = int((4 * 24 * (lists + 1)).sum())
        intra = int(((p.mb_info["mb_class"] >= abi.MB_I4x4) & (p.mb_info["mb_class"] <= abi.MB_IPCM)).sum()) * 384
        dbk = 2 * rp.frame_bytes if p.deblock_enable else 0
        out.append({"inter": mc, "intra": intra, "deblock": dbk})
    return out


def measure_config(name, stems_, S, local_rank, world, max_pictures=None):
    """Short device-timed measurement of another BASELINE.json configuration (one warm-up pass that is checked against the reference's
    checksums, one timed pass): value in frames/s over all ranks, and whether parity was checked on full streams."""
    import numpy as np
    from h264_video_decoder_demo_b200 import engine, replay, sharding
    eng, vs, full_all = None, [], True
    for st in stems_:
        path, full = find_replay(st)
        full_all &= full
        rpv = replay.load_replay(path) if max_pictures is None else replay.parse_replay(replay.read_replay_bytes(path), path, max_pictures)
        if eng is None:
            eng = engine.Engine(local_rank, S, rpv.width_mbs, rpv.height_mbs)
        vs.append({"rp": rpv, "rs": None})
    sids = list(range(S))
    var_of = [s % len(vs) for s in sids]
    rs = []
    for s in sids:
        v = vs[var_of[s]]
        if v["rs"] is None:
            v["rs"] = engine.ResidentStream(eng, v["rp"]); rs.append(v["rs"])
        else:
            rs.append(v["rs"].clone())
    npic = max(len(v["rp"].pictures) for v in vs)
    pic_of = lambda s, i: i % len(vs[var_of[s]]["rp"].pictures)
    batches = [eng.prepare(sids, [rs[s].params[pic_of(s, i)] for s in sids]) for i in range(npic)]
    dst = [[vs[var_of[s]]["rp"].pictures[pic_of(s, i)].dst_surface for s in sids] for i in range(npic)]
    want = [[vs[var_of[s]]["rp"].pictures[pic_of(s, i)].sum_post for s in sids] for i in range(npic)]
    ok = True
    for i, b in enumerate(batches):                 # warm-up pass, every picture of every stream checked
        eng.submit_prepared(b)
        ok &= eng.checksums(sids, dst[i]) == want[i]
    eng.sync()
    sharding.barrier()
    eng.timer_start()
    for b in batches:
        eng.submit_prepared(b)
    ms = sharding.max_over_ranks(eng.timer_stop())
    kt = eng.kernel_times()
    for r_ in rs:
        r_.free()
    eng.close()
    if not ok:
        raise SystemExit(f"PARITY FAILURE in configuration {name}")
    return {"value": round(S * npic * world / (ms / 1000.0), 1), "unit": UNIT, "streams_per_gpu": S, "pictures_per_pass_per_gpu": S * npic,
            "ms_per_pass": round(ms, 1), "parity": "every picture of every stream equals the reference's checksum" + ("" if full_all else " (golden prefixes)"),
            "kernel_ms": {k: round(v["ms"], 1) for k, v in kt.items()}}


_T0 = time.time()


def phase(msg):
    """Progress line on stderr (rank 0 only; stdout carries nothing but the one JSON line)."""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write("bench [%6.1f s] %s\n" % (time.time() - _T0, msg))
        sys.stderr.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=0, help="concurrent replicas of the stream per GPU (0 = 384, halved until the replicas fit in HBM)")
    ap.add_argument("--workload", default="B_frames.cabac", choices=sorted(WORKLOADS) + ["mixed"],
                    help="one bundled stream replicated S times, or 'mixed' = all five bundled variants dealt round-robin over the S streams")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", default="full", choices=["full", "h2d", "d2h"], help="diagnostic: which PCIe legs the e2e loop includes")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--host-layout", default="shared", choices=["batch", "shared"], help="e2e: picture blocks of a submit back to back in pinned memory (one DMA per batch) or one block per distinct picture")
    ap.add_argument("--dense-coefs", action="store_true", help="e2e: send plain arrays (dense int16 levels, full motion records) instead of the packed transport")
    ap.add_argument("--no-pack-in-e2e", action="store_true", help="e2e: pack levels / motion once before the timed region instead of for every submitted picture inside it")
    ap.add_argument("--no-configs", action="store_true", help="skip the short measurements of the other BASELINE.json configurations (tff, gop121, mixed)")
    ap.add_argument("--no-bitstream", action="store_true", help="skip the Annex-B-in / frames-out pipeline measurement")
    ap.add_argument("--bitstream-streams", type=int, default=32, help="streams per GPU of the Annex-B pipeline measurement")
    ap.add_argument("--bitstream-threads", type=int, default=0, help="parser threads (0 = usable host cores / ranks)")
    ap.add_argument("--max-pictures", type=int, default=None)
    args = ap.parse_args()
    stems = [WORKLOADS[w] for w in (MIXED if args.workload == "mixed" else [args.workload])]
    stem = stems[0]

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        import socket
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        reference_arm(args, stem)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) must not write to stdout: the contract is ONE JSON line there
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    from h264_video_decoder_demo_b200 import engine, replay, sharding
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    S = args.streams or 384
    eng_probe = engine.load_library()   # fails loudly when the CUDA library is missing: no fallback
    del eng_probe
    rs, eng = [], None
    while True:
        try:
            # one variant per distinct bitstream: page-locked copy of its container (the e2e leg DMAs every picture's arrays
            # straight from it) + replicas resident in HBM, one distinct copy of the SoA per stream
            variants, is_full, eng = [], True, None
            for st in stems:
                path, full = find_replay(st)
                is_full &= full
                raw = replay.read_replay_bytes(path)
                if eng is None:
                    rp0 = replay.parse_replay(raw, path, 1)
                    eng = engine.Engine(local_rank, S, rp0.width_mbs, rp0.height_mbs)
                pinned = eng.pinned_array(len(raw))
                pinned[:] = np.frombuffer(raw, dtype=np.uint8)
                del raw
                rpv = replay.parse_replay(pinned, path, args.max_pictures)
                # what the host stage hands to h264b2_submit: one page-locked block per picture (the layout of the front end's picture
                # blocks: arrays back to back, 64-byte aligned -> ONE DMA per picture), coefficients packed by h264b2_pack_coefs
                blobs, host_params = None, None
                if not args.dense_coefs:
                    names = ("mb_info", "intra_modes", "coef_offset", "motion", "weights", "level_scale4", "level_scale8")
                    al = lambda n: (n + 63) & ~63
                    cb = lambda pic: int(eng.lib.h264b2_pack_coefs_bound(len(pic.coefs)))
                    mb_ = lambda pic: int(eng.lib.h264b2_pack_coefs_bound(pic.motion.size * 76)) if pic.motion is not None else 0
                    bound = sum(sum(al(getattr(pic, k).nbytes) for k in names if getattr(pic, k) is not None) + al(cb(pic)) + al(mb_(pic)) + 64 for pic in rpv.pictures)
                    store, blobs, mblobs, host_params, o = eng.pinned_array(bound), [], [], [], 0
                    o += (-store.ctypes.data) % 64
                    extents = []
                    t_pack = 0.0
                    for pic in rpv.pictures:
                        ptrs, o0, mblob = {}, o, None
                        for k in names:
                            a = getattr(pic, k)
                            if a is None or not a.size:
                                continue
                            if k == "motion":
                                if not pic.has_inter:
                                    continue
                                t0p = time.perf_counter()
                                mblob = engine.pack_motion(a, store[o:o + mb_(pic)])
                                t_pack += time.perf_counter() - t0p
                                ptrs[k] = mblob.ctypes.data
                                o += al(mblob.size)
                                continue
                            store[o:o + a.nbytes] = a.view(np.uint8).reshape(-1)
                            ptrs[k] = store.ctypes.data + o
                            o += al(a.nbytes)
                        t0p = time.perf_counter()
                        b = engine.pack_coefs(pic.coefs, store[o:o + cb(pic)])
                        t_pack += time.perf_counter() - t0p
                        o += al(b.size)
                        blobs.append(b)
                        mblobs.append(mblob)
                        ptrs["coefs"] = b.ctypes.data
                        host_params.append(replay.pic_params(rpv, pic, ptrs=ptrs, packed_blob=b, packed_motion=mblob))
                        extents.append((o0, o - o0, ptrs))
                else:
                    host_params = [replay.pic_params(rpv, pic) for pic in rpv.pictures]
                variants.append({"rp": rpv, "host_params": host_params, "_store": store if blobs else None, "extents": extents if blobs else None, "blobs": blobs, "mblobs": mblobs if blobs else None,
                                 "pack_ms": (1000.0 * t_pack / max(1, len(rpv.pictures))) if blobs else None,
                                 "host_bytes": [e[1] for e in extents] if blobs else [pic.nbytes() for pic in rpv.pictures], "rs": None})
            sids = list(range(S))
            var_of = [s % len(variants) for s in sids]
            rs = []
            for s in sids:
                v = variants[var_of[s]]
                if v["rs"] is None:
                    v["rs"] = engine.ResidentStream(eng, v["rp"])
                    rs.append(v["rs"])
                else:
                    rs.append(v["rs"].clone())

            break
        except engine.EngineError as ex:
            # the replicas (one distinct SoA copy + a 17-surface DPB per stream) did not fit: halve the stream count (only when it was not given)
            if args.streams or S <= 32 or "memory" not in str(ex).lower():
                raise
            sys.stderr.write(f"bench: {S} streams do not fit in device memory ({ex}); retrying with {S // 2}\n")
            try:
                for r_ in rs:
                    r_.free()
                if eng is not None:
                    eng.close()
            except Exception:
                pass
            S //= 2
    rp = variants[0]["rp"]
    npic = max(len(v["rp"].pictures) for v in variants)          # submits per step; shorter streams wrap to their IDR

    def pic_of(s, i):
        return i % len(variants[var_of[s]]["rp"].pictures)

    batches = [eng.prepare(sids, [rs[s].params[pic_of(s, i)] for s in sids]) for i in range(npic)]
    # Host layout of the e2e leg.  "batch": the picture blocks of one submit lie back to back in page-locked memory (a batch arena,
    # as the host stage's allocator hands them out), so h264b2_submit moves a batch in ONE DMA and the read-back of a batch is
    # one DMA too — picture-sized transfers cost 35 % of the link when both directions run (profiles/r01_pcie_probe2.txt).
    # "shared": one block per distinct picture, one DMA per picture (also the fallback when host memory is short).
    host_layout, arena = "shared", None
    if not args.dense_coefs and not args.no_e2e and args.host_layout == "batch":
        need = sum(variants[var_of[s]]["extents"][pic_of(s, i)][1] for s in sids for i in range(npic)) + 4096
        avail = int(re.search(r"MemAvailable:\s+(\d+)", open("/proc/meminfo").read()).group(1)) * 1024
        if avail > 3 * need * world:          # every rank of the node allocates the same arena at the same moment
            host_layout = "batch"
            arena = eng.pinned_array(need)
            o = (-arena.ctypes.data) % 64
            host_batches = []
            for i in range(npic):
                plist = []
                for s in sids:
                    v = variants[var_of[s]]
                    k = pic_of(s, i)
                    o0, ln, ptrs = v["extents"][k]
                    arena[o:o + ln] = v["_store"][o0:o0 + ln]
                    delta = arena.ctypes.data + o - (v["_store"].ctypes.data + o0)
                    blob, mblob = v["blobs"][k], v["mblobs"][k]
                    pp = replay.pic_params(v["rp"], v["rp"].pictures[k], ptrs={n_: a_ + delta for n_, a_ in ptrs.items()}, packed_blob=blob, packed_motion=mblob)
                    if pp.packed & 1:
                        pp.coefs = blob.ctypes.data + delta
                    if pp.packed & 2:
                        pp.motion = mblob.ctypes.data + delta
                    plist.append(pp)
                    o += ln
                host_batches.append(eng.prepare(sids, plist))
    if host_layout == "shared":
        host_batches = [eng.prepare(sids, [variants[var_of[s]]["host_params"][pic_of(s, i)] for s in sids]) for i in range(npic)]
    dst = [[variants[var_of[s]]["rp"].pictures[pic_of(s, i)].dst_surface for s in sids] for i in range(npic)]
    want = [[variants[var_of[s]]["rp"].pictures[pic_of(s, i)].sum_post for s in sids] for i in range(npic)]

    def step():
        for b in batches:
            eng.submit_prepared(b)

    phase("%d replicas resident, batches prepared" % S)
    # ---- warm-up (the last warm-up step is checked against the reference's checksums, every stream, every picture)
    for w in range(max(args.warmup, 1)):
        if w == max(args.warmup, 1) - 1:
            for i, b in enumerate(batches):
                eng.submit_prepared(b)
                sums = eng.checksums(sids, dst[i])
                if sums != want[i]:
                    raise SystemExit(f"PARITY FAILURE: submit {i}: {sum(x != y for x, y in zip(sums, want[i]))} of {S} streams differ from the reference")
        else:
            step()
    eng.sync()

    phase("warm-up done, parity checked against the reference's checksums")
    # ---- timed region: device time by CUDA events on the launch stream, max over ranks
    sampler = ClockSampler(local_rank)
    time.sleep(0.3)
    sharding.barrier()
    eng.sync()
    t_wall0 = time.time()
    eng.timer_start()
    for _ in range(args.steps):
        step()
    ms = eng.timer_stop()
    t_wall1 = time.time()
    sharding.barrier()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_max = sharding.max_over_ranks(ms)
    # per-kernel durations for the roofline: one more step with the look-ahead stream off, i.e. every kernel alone on the launch
    # stream (with it on, k_residual / k_bs of the next batch overlap the wavefront kernels and event-bracketed durations
    # would charge each kernel for its neighbours)
    eng.set_lookahead(False)
    step()
    eng.sync()
    eng.timer_start()
    step()
    ms_serial = eng.timer_stop()
    kt = eng.kernel_times()
    eng.set_lookahead(True)
    kt_steps = 1
    pictures_per_rank = S * npic * args.steps
    value = pictures_per_rank * world / (ms_max / 1000.0)
    # final state check: the last picture of every stream still equals the reference
    if eng.checksums(sids, dst[-1]) != want[-1]:
        raise SystemExit("PARITY FAILURE after the timed region")

    phase("timed steps done")
    # ---- roofline of the dominant kernel
    abv = [algorithmic_bytes(v["rp"]) for v in variants]
    per_step = {k: sum(abv[var_of[s]][pic_of(s, i)][k] for s in sids for i in range(npic)) for k in ("inter", "intra", "deblock")}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kernels = {}
    for k in ("inter", "intra", "deblock"):
        t = kt[k]["ms"] + (kt["bs"]["ms"] if k == "deblock" else 0.0)
        ach = per_step[k] * kt_steps / (t / 1000.0) / 1e9 if t > 0 else None
        kernels[k] = {"ms_per_step": round(t / kt_steps, 3), "launches_per_step": kt[k]["launches"] // kt_steps,
                      "algorithmic_gb_per_step": round(per_step[k] / 1e9, 4),
                      "achieved_gbs": round(ach, 1) if ach else None, "frac": round(ach / peak, 4) if ach else None}
    # The roofline object is the one BASELINE.json's north star asks for: motion compensation + deblocking (k_inter_tma + k_inter_list,
    # k_bs_prog2 + k_deblock3) against the measured HBM copy bandwidth.  The largest kernel by time is reported next to it.
    dom_by_time = max(("inter", "intra", "deblock"), key=lambda k: kernels[k]["ms_per_step"])
    dom = "deblock" if kernels["deblock"]["ms_per_step"] >= kernels["inter"]["ms_per_step"] else "inter"
    mc_db_ms = kernels["inter"]["ms_per_step"] + kernels["deblock"]["ms_per_step"]
    mc_db_bytes = per_step["inter"] + per_step["deblock"]
    achieved = round(mc_db_bytes / (mc_db_ms / 1000.0) / 1e9, 1) if mc_db_ms > 0 else None
    dom_ms = mc_db_ms
    # DRAM traffic from the ncu --set full capture of THIS build (tools/gpu_traffic.sh + tools/make_traffic.py write
    # profiles/traffic_r02.json with the SHA-1 of the engine sources); a capture of other sources is not reported
    traffic, traffic_src = None, "no ncu capture of these sources (profiles/traffic_r02.json absent or of another source state)"
    try:
        import hashlib
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
        from h264_video_decoder_demo_b200 import build as _b
        if tr.get("src_sha1") == _b.source_hash() and "H264B2_LIB" not in os.environ:
            per_pic = sum((tr.get(k) or {}).get("dram_bytes_per_picture") or 0 for k in ("k_inter", "k_deblock", "k_bs"))
            traffic = int(per_pic * S) if per_pic else None      # per launch = per picture x pictures per launch
            traffic_src = "ncu dram__bytes_read.sum + dram__bytes_write.sum of this build (profiles/traffic_r02.json): MC + bS + deblock kernels, per batch of %d pictures" % S
            for k, n in (("inter", "k_inter"), ("deblock", "k_deblock")):
                if tr.get(n):
                    kernels[k]["dram_bytes_per_picture"] = tr[n]["dram_bytes_per_picture"]; kernels[k]["algorithmic_bytes_per_picture_of_capture"] = tr[n]["algorithmic_bytes_per_picture"]
                    kernels[k]["warp_instructions_per_picture"] = tr[n]["warp_instructions_per_picture"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "MC + deblocking: k_inter_tma + k_inter_list, k_bs_prog2 + k_deblock3", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "share_of_step": round(dom_ms / ms_serial, 3), "largest_kernel_by_time": {"inter": "k_inter_tma+k_inter_list", "intra": "k_intra", "deblock": "k_bs_prog2+k_deblock3"}[dom_by_time],
                "largest_of_mc_deblock": {"inter": "k_inter_tma+k_inter_list", "deblock": "k_bs_prog2+k_deblock3"}[dom], "kernels": kernels,
                "timing": "CUDA events around every launch on its stream, one extra step with the look-ahead stream off (kernels serialised): %.1f ms; the timed steps overlap k_residual/k_bs of batch i+1 with the wavefront kernels of batch i: %.1f ms per step" % (ms_serial, ms / args.steps)}
    kernels["residual"] = {"ms_per_step": round(kt["residual"]["ms"] / kt_steps, 3), "launches_per_step": kt["residual"]["launches"] // kt_steps}
    kernels["deblock"]["of_which_bs_ms"] = round(kt["bs"]["ms"] / kt_steps, 3)
    launches = sum(kt[k]["launches"] for k in ("residual", "inter", "intra", "bs", "deblock")) * args.steps

    phase("per-kernel timing done")
    # ---- e2e: host buffers in, host pictures out, through the C ABI
    e2e = None
    if not args.no_e2e:
        nbuf = 2
        out_host = [eng.pinned_array(S * eng.frame_bytes) for _ in range(nbuf)]
        out_ptrs = [[int(o.ctypes.data) + s * eng.frame_bytes for s in sids] for o in out_host]

        # Host packing INSIDE the timed region: the host stage packs the levels and motion records of every (stream, picture) it
        # submits (h264b2_pack_coefs / h264b2_pack_motion).  The replicas share one page-locked copy of each packed picture, so the
        # work is done for real — S x pictures pack calls per step on a pool of host threads, one batch ahead of its submit — and
        # its output goes to per-thread scratch buffers (the bytes are identical to the shared copy the DMA reads).
        pack_pool, pack_futs, n_pack_threads = None, {}, 0
        if not args.dense_coefs and not args.no_pack_in_e2e:
            from concurrent.futures import ThreadPoolExecutor
            n_pack_threads = max(1, usable_cores() // world - 1)
            pack_pool = ThreadPoolExecutor(max_workers=n_pack_threads)
            scratch = {}

            def pack_range(i, lo, hi):
                import threading as _th
                tid = _th.get_ident()
                for s_ in range(lo, hi):
                    v = variants[var_of[s_]]
                    pic = v["rp"].pictures[pic_of(s_, i)]
                    cbn = int(eng.lib.h264b2_pack_coefs_bound(len(pic.coefs)))
                    mbn = int(eng.lib.h264b2_pack_coefs_bound(pic.motion.size * 76)) if pic.motion is not None and pic.has_inter else 0
                    buf = scratch.get(tid)
                    if buf is None or buf.size < cbn + mbn + 64:
                        raw_ = np.empty(max(cbn + mbn + 64, 8 << 20), np.uint8)
                        buf = scratch[tid] = raw_[(-raw_.ctypes.data) % 16:]
                    engine.pack_coefs(pic.coefs, buf[:cbn])
                    if mbn:
                        o_ = (cbn + 15) & ~15
                        engine.pack_motion(pic.motion, buf[o_:o_ + mbn])

            def pack_batch(i):
                per = (S + n_pack_threads - 1) // n_pack_threads
                return [pack_pool.submit(pack_range, i, lo, min(S, lo + per)) for lo in range(0, S, per)]

        def e2e_step():
            if pack_pool is not None:
                pack_futs[0] = pack_batch(0)
            for i, b in enumerate(host_batches):
                if pack_pool is not None:
                    if i + 1 < npic:
                        pack_futs[i + 1] = pack_batch(i + 1)
                    for f_ in pack_futs.pop(i):
                        f_.result()
                if args.e2e_mode == "d2h":
                    eng.submit_prepared(batches[i])
                else:
                    eng.submit_prepared_host(b)
                if args.e2e_mode != "h2d":
                    eng.read_pictures_async(sids, dst[i], out_ptrs[i % nbuf])

        e2e_step()
        eng.sync()
        # the frames that came back are the reference's frames
        last = (npic - 1) % nbuf
        from h264_video_decoder_demo_b200 import abi
        for s in ((0, S - 1) if args.e2e_mode != "h2d" else ()):
            if abi.checksum(out_host[last][s * eng.frame_bytes:(s + 1) * eng.frame_bytes].tobytes()) != want[-1][s]:
                raise SystemExit("PARITY FAILURE in the e2e path")
        sharding.barrier()
        eng.timer_start()
        t0 = time.time()
        for _ in range(args.e2e_steps):
            e2e_step()
        eng.sync()
        t1 = sharding.max_over_ranks(time.time() - t0)
        e2e_dev_ms = eng.timer_stop()
        e2e_kt = eng.kernel_times()
        e2e = {"value": round(S * npic * args.e2e_steps * world / t1, 1), "unit": UNIT,
               "h2d_bytes_per_step": int(sum(variants[var_of[s]]["host_bytes"][pic_of(s, i)] for s in sids for i in range(npic))),
               "host_pack_ms_per_picture": None if args.dense_coefs else round(max(v["pack_ms"] for v in variants), 3),
               "host_pack_in_timed_region": pack_pool is not None, "host_pack_threads": n_pack_threads,
               "host_pack_note": ("h264b2_pack_coefs + h264b2_pack_motion run INSIDE the timed region for every (stream, picture) of every step, on %d host threads, one batch ahead of its submit" % n_pack_threads) if pack_pool is not None
                                 else "h264b2_pack_coefs + h264b2_pack_motion done once per picture BEFORE the timed region (the timed region starts at the C-ABI call)",
               "host_layout": host_layout, "arrays": "plain (dense int16 levels, 152-byte motion records)" if args.dense_coefs else "packed levels and motion records (h264b2_pack_coefs / h264b2_pack_motion)",
               "d2h_bytes_per_step": int(npic * S * eng.frame_bytes),
               "steps": args.e2e_steps, "device_ms_per_step": round(e2e_dev_ms / args.e2e_steps, 1),
               "kernel_ms_per_step": {k: round(v["ms"] / args.e2e_steps, 1) for k, v in e2e_kt.items()}, "timing": "host wall clock around submit+read-back of every picture, synchronised on both sides, max over ranks"}

    l2_gb = (sum(sum(r.blob_bytes) for r in rs) + 17 * eng.frame_bytes * S) / 1e9
    phase("e2e leg done")
    # ---- the whole decoder: Annex-B in, pictures out (host entropy/derivation stage on a thread pool + the CUDA engine)
    e2e_bits = None
    stream_files = [os.path.join(REF_DIR, "streams", st + ".h264") for st in stems]
    if not args.no_bitstream and all(os.path.exists(f) for f in stream_files):
        from h264_video_decoder_demo_b200 import frontend
        for r_ in rs:
            r_.free()
        rs = []
        eng.close()
        nstr = args.bitstream_streams
        threads = args.bitstream_threads or max(1, usable_cores() // world)
        paths = [stream_files[i % len(stream_files)] for i in range(nstr)]
        hashes_want = [frontend.hash_chain(variants[i % len(variants)]["rp"].out_sums) for i in range(nstr)] if is_full and args.max_pictures is None else None
        frontend.multi_decode(paths[:min(nstr, 4)], device=local_rank, threads=min(threads, 4), readback=True, hashes=False)      # warm-up (page-locked pools, clocks)
        sharding.barrier()
        st_b, hs = frontend.multi_decode(paths, device=local_rank, threads=threads, readback=True, hashes=True)
        if hashes_want is not None and hs != hashes_want:
            raise SystemExit("PARITY FAILURE in the Annex-B pipeline: output frames differ from the reference decoder's")
        secs = sharding.max_over_ranks(st_b["seconds"])
        e2e_bits = {"value": round(st_b["frames_out"] * world / secs, 1), "unit": UNIT, "streams_per_gpu": nstr, "parser_threads_per_gpu": st_b["threads"],
                    "host_cores": usable_cores(), "frames": st_b["frames_out"] * world, "seconds": round(secs, 3),
                    "host_stage_pictures_per_s_per_thread": round(st_b["pictures"] / st_b["parse_seconds"], 1) if st_b["parse_seconds"] > 0 else None,
                    "h2d_bytes": st_b["h2d_bytes"], "d2h_bytes": st_b["d2h_bytes"], "submits": st_b["submits"],
                    "checked": "every stream's output-order checksum chain equals the reference decoder's" if hashes_want is not None else "not checked (prefix fixtures)",
                    "what": "h264b2_multi_decode: Annex-B byte streams in, every output picture in page-locked host memory; host entropy decoding + derivations included"}

    phase("whole-decoder leg done")
    # ---- the other BASELINE.json configurations, shortly (device-resident replay, parity-checked): config 3 (field/MBAFF stream), config 4
    #      (the long-GOP stream; it has ONE closed GOP, so GOP-parallelism is shown on the two-GOP HeavyHand stream by the tests and the
    #      multi-stream pipeline), config 5 (64 streams of all five variants, STRONG scaling: 64 / N streams per GPU)
    configs = None
    if not args.no_configs and args.workload == "B_frames.cabac" and args.max_pictures is None:
        try:
            for r_ in rs:
                r_.free()
            rs = []
            eng.close()
        except Exception:
            pass
        configs = {}
        configs["tff_mbaff"] = dict(measure_config("tff", [WORKLOADS["tff"]], min(S, 128), local_rank, world), baseline_config=3, scaling="weak")
        configs["gop121_long_gop"] = dict(measure_config("gop121", [WORKLOADS["gop121"]], min(S, 128), local_rank, world), baseline_config=4, scaling="weak",
                                          note="one closed GOP (single IDR): replicas only; closed-GOP splitting is exercised on the HeavyHand streams (tests, H264B2_MULTI_SPLIT_GOPS)")
        per_gpu = max(1, 64 // world)
        configs["mixed_64_streams"] = dict(measure_config("mixed", [WORKLOADS[w] for w in MIXED], per_gpu, local_rank, world), baseline_config=5, scaling="strong",
                                           note="64 streams in total, all five bundled variants round-robin, 64 / n_gpus per GPU")

    phase("configs block done")
    # ---- CPU baseline: the unmodified reference on one host core (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = run_reference_once(stem, 16, 1)
        if r:
            cpu = {"value": round(r[0] / r[1], 3), "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"oracle/_ref/ref_harness (unmodified reference, its Makefile flags) on {stem}.h264 until 16 frames reached the callback: {r[0]} pictures in {r[1]:.1f} s"}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref/ref_harness not present on this box"}

    if rank == 0:
        os.write(real_stdout, (json.dumps({
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "bundled reference bitstream pre-parsed by the reference's own parser (SoA resident in HBM)",
            "config": {"workload": (f"{stem}.h264 x {S} concurrent replicas per GPU" if len(stems) == 1 else
                                    f"all {len(stems)} bundled streams ({', '.join(stems)}) dealt round-robin over {S} concurrent streams per GPU")
                                   + ("" if is_full else " (golden prefix only: full replay not built)"),
                       "pictures_per_step_per_gpu": S * npic, "streams_per_gpu": S, "picture": "1920x1088 I420",
                       "l2": "inputs larger than L2: one distinct SoA copy + 17-surface DPB per stream (%.1f GB per GPU)" % l2_gb,
                       "parallelism": f"streams sharded over {world} GPU(s), no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_bitstream": e2e_bits,
            # the ratio-ready whole-decoder number (Annex-B in, frames out, host entropy decoding included): like for like with the reference arm
            "e2e_whole_decoder": ({"value": e2e_bits["value"], "unit": UNIT, "what": "h264b2_multi_decode (Annex-B byte streams in, frames in host memory out); compare THIS with the reference arm, which decodes from the bitstream too"} if e2e_bits else None),
            "configs": configs, "gpu_launches": launches, "clocks": clocks,
        }) + "\n").encode())
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
