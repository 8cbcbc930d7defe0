"""Turn an `ncu --set full` report of the first pass of `bench.py --streams 128 --max-pictures 10` (tools/gpu_traffic.sh) into
profiles/traffic_r02.json: DRAM bytes per launch / per picture of every hot kernel NEXT TO the algorithmic bytes of exactly the pictures
those launches reconstructed, and the SHA-1 of the engine sources that were profiled (h264_video_decoder_demo_b200/build.py: source_hash; bench.py reports `roofline.traffic` only for that source state — the compiled library is not bit-reproducible).
usage: python tools/make_traffic.py gpurun_out/prof_r02.ncu-rep gpurun_out/prof_r02.libsha1 [streams]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(rep, shafile, streams=128):
    import bench
    from h264_video_decoder_demo_b200 import replay
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, data = rows[0], rows[2:]
    ki, ri, wi, ti, ii = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("gpu__time_duration.sum"), H.index("smsp__inst_executed.sum")
    ur, uw = rows[1][ri], rows[1][wi]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    path, _ = bench.find_replay(bench.WORKLOADS["B_frames.cabac"])
    rp = replay.parse_replay(replay.read_replay_bytes(path), path, 10)
    ab = bench.algorithmic_bytes(rp)
    # launches of a kernel, in order <-> the pictures it is launched for, in order
    launched = {"k_inter_tma": [i for i, p in enumerate(rp.pictures) if p.has_inter], "k_inter_list": [i for i, p in enumerate(rp.pictures) if p.has_inter],
                "k_deblock3": [i for i, p in enumerate(rp.pictures) if p.deblock_enable], "k_bs_prog2": [i for i, p in enumerate(rp.pictures) if p.deblock_enable],
                "k_intra": list(range(len(rp.pictures))), "k_residual": list(range(len(rp.pictures)))}
    alg_key = {"k_inter_tma": "inter", "k_inter_list": "inter", "k_deblock3": "deblock", "k_bs_prog2": "deblock", "k_intra": "intra", "k_residual": None}
    res = {"source": "ncu --set full --clock-control none over the first pass of bench.py --streams %d --max-pictures 10 (tools/gpu_traffic.sh)" % streams,
           "src_sha1": open(shafile).read().split()[0], "streams_per_launch": streams, "kernels": {}}
    seen = {}
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        if name not in launched:
            continue
        k = seen.get(name, 0)
        seen[name] = k + 1
        if k >= len(launched[name]):
            continue                                   # second pass: same pictures again
        pic = launched[name][k]
        e = res["kernels"].setdefault(name, {"launches": []})
        dram = float(r[ri].replace(",", "")) * scale.get(ur, 1.0) + float(r[wi].replace(",", "")) * scale.get(uw, 1.0)
        e["launches"].append({"picture": pic, "dram_bytes": int(dram), "algorithmic_bytes": (ab[pic][alg_key[name]] * streams) if alg_key[name] else None,
                              "duration_ms": float(r[ti].replace(",", "")), "warp_instructions": float(r[ii].replace(",", ""))})
    for name, e in res["kernels"].items():
        L = e["launches"]
        e["dram_bytes_per_launch"] = int(sum(x["dram_bytes"] for x in L) / len(L))
        e["dram_bytes_per_picture"] = int(e["dram_bytes_per_launch"] / streams)
        e["warp_instructions_per_picture"] = int(sum(x["warp_instructions"] for x in L) / len(L) / streams)
        if alg_key[name]:
            a = sum(x["algorithmic_bytes"] for x in L) / len(L)
            e["algorithmic_bytes_per_launch"] = int(a)
    K = res["kernels"]
    # what bench.py looks up: the three roofline classes (inter = staged + list kernel, deblock = filter kernel)
    def cls(names, algname):
        n = [K[x] for x in names if x in K]
        if not n:
            return None
        d = sum(x["dram_bytes_per_picture"] for x in n)
        a = n[0].get("algorithmic_bytes_per_launch")
        return {"dram_bytes_per_picture": d, "algorithmic_bytes_per_picture": int(a / streams) if a else None, "ratio": round(d / (a / streams), 3) if a else None,
                "warp_instructions_per_picture": sum(x["warp_instructions_per_picture"] for x in n)}
    res["k_inter"] = cls(["k_inter_tma", "k_inter_list"], "inter")
    res["k_deblock"] = cls(["k_deblock3"], "deblock")
    res["k_bs"] = cls(["k_bs_prog2"], "deblock")
    res["k_intra"] = cls(["k_intra"], "intra")
    json.dump(res, open(os.path.join(ROOT, "profiles", "traffic_r02.json"), "w"), indent=1)
    print(json.dumps({k: res[k] for k in ("k_inter", "k_deblock", "k_bs", "k_intra")}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 128)
