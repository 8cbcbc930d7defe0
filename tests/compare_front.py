"""Compare a picture container written by the native front end (tools/h264b2_parse) with the one the reference's own
parser produced (oracle/ref_harness --replay): every field of every picture, in decoding order.  Test infrastructure."""
import sys

import numpy as np


def compare(mine, ref, max_report=5, strict_mv=True):
    """Returns a list of human-readable differences (empty = identical up to the checksum fields)."""
    diffs = []
    if (mine.width_mbs, mine.height_mbs) != (ref.width_mbs, ref.height_mbs):
        return [f"size {mine.width_mbs}x{mine.height_mbs} != {ref.width_mbs}x{ref.height_mbs}"]
    if len(mine.pictures) != len(ref.pictures):
        diffs.append(f"picture count {len(mine.pictures)} != {len(ref.pictures)}")
    for i, (a, b) in enumerate(zip(mine.pictures, ref.pictures)):
        d = []
        for f in ("decode_idx", "dst_surface", "clear_surface", "has_inter", "deblock_enable", "deblock_stop_mb", "mbaff", "cqp", "slice_type", "poc", "n_na", "nal_ref_idc"):
            if getattr(a, f) != getattr(b, f):
                d.append(f"{f}: {getattr(a, f)} != {getattr(b, f)}")
        for f in a.mb_info.dtype.names:
            if f == "reserved":
                continue
            ne = np.nonzero(a.mb_info[f] != b.mb_info[f])[0]
            if ne.size:
                m = int(ne[0])
                d.append(f"mb_info.{f}: {ne.size} MBs differ, first MB {m}: {a.mb_info[f][m]} != {b.mb_info[f][m]}")
        ne = np.nonzero(a.intra_modes != b.intra_modes)[0]
        if ne.size:
            m = int(ne[0]); d.append(f"intra_modes: {ne.size} MBs differ, first MB {m}: {int(a.intra_modes[m]):#x} != {int(b.intra_modes[m]):#x}")
        if len(a.coefs) != len(b.coefs) or not np.array_equal(a.coefs, b.coefs) or not np.array_equal(a.coef_offset, b.coef_offset):
            ne = np.nonzero(a.coef_offset != b.coef_offset)[0]
            first = int(ne[0]) - 1 if ne.size else -1
            if first < 0 and len(a.coefs) == len(b.coefs):
                pos = int(np.nonzero(a.coefs != b.coefs)[0][0]); first = int(np.searchsorted(b.coef_offset, pos, side="right")) - 1
            d.append(f"coefs: {len(a.coefs)} vs {len(b.coefs)} values, first differing MB ~{first}")
        if (a.motion is None) != (b.motion is None):
            d.append("motion presence differs")
        elif a.motion is not None:
            inter = b.mb_info["mb_class"] == 5
            for f in ("ref_surf", "ref_ident", "wt_idx"):
                ne = np.nonzero((a.motion[f] != b.motion[f]).reshape(len(inter), -1).any(axis=1) & inter)[0]
                if ne.size:
                    m = int(ne[0]); d.append(f"motion.{f}: {ne.size} MBs differ, first MB {m}: {a.motion[f][m].tolist()} != {b.motion[f][m].tolist()}")
            mva, mvb = a.motion["mv"], b.motion["mv"]            # [n][2][16][2]
            if strict_mv:
                bad = (mva != mvb).reshape(len(inter), -1).any(axis=1) & inter
            else:                                                 # only lists that are used (ref_surf >= 0) in the block's quadrant
                quad = np.array([(r // 2) * 2 + (c // 2) for r in range(4) for c in range(4)])
                used = (b.motion["ref_surf"][:, :, quad] >= 0)  # [n][2][16]
                bad = ((mva != mvb).any(axis=3) & used).reshape(len(inter), -1).any(axis=1) & inter
            ne = np.nonzero(bad)[0]
            if ne.size:
                m = int(ne[0]); d.append(f"motion.mv: {ne.size} MBs differ, first MB {m}: {mva[m].tolist()} != {mvb[m].tolist()}")
        if len(a.weights) != len(b.weights) or a.weights.tobytes() != b.weights.tobytes():
            d.append(f"weights: {len(a.weights)} vs {len(b.weights)} entries / content differs")
        for f in ("level_scale4", "level_scale8"):
            x, y = getattr(a, f), getattr(b, f)
            if (x is None) != (y is None) or (x is not None and not np.array_equal(x, y)):
                d.append(f"{f} differs")
        if d:
            diffs.append(f"picture {i}: " + "; ".join(d))
            if len(diffs) >= max_report:
                break
    if mine.out_order != ref.out_order:
        diffs.append(f"output order differs: {mine.out_order[:12]}... != {ref.out_order[:12]}...")
    return diffs


if __name__ == "__main__":
    import os
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    from h264_video_decoder_demo_b200 import replay
    mine, ref = replay.load_replay(sys.argv[1]), replay.load_replay(sys.argv[2])
    if len(sys.argv) > 3:
        n = int(sys.argv[3]); ref.pictures = ref.pictures[:n]; mine.pictures = mine.pictures[:n]
        ref.out_order = [x for x in ref.out_order if x < n]; mine.out_order = [x for x in mine.out_order if x < n]
    out = compare(mine, ref)
    print("\n".join(out) if out else f"identical: {len(ref.pictures)} pictures")
    sys.exit(1 if out else 0)
