"""A small H.264 *syntax* writer (CAVLC, progressive, I/P slices) for test streams.  TEST INFRASTRUCTURE.

It does not encode video: it draws random but well-formed syntax elements (macroblock types incl. P sub-partitions 8x4/4x8/4x4,
P_8x8ref0, Intra16x16 with every prediction mode, Intra4x4/8x8 with predicted and explicit modes, I_PCM, multiple slices per
picture with their own deblocking parameters, several reference pictures, explicit weighted prediction, 8x8 transform) and
serialises them as an Annex-B byte stream.  The UNMODIFIED reference decoder (oracle/_ref/ref_harness) then decodes the stream and
its interpretation — whatever it is — becomes the golden answer the native front end and the CUDA engine must reproduce
(tools/make_golden.py writes the fixtures).  The bundled streams never reach most of these paths.

VLC tables are read from the product's generated table file (h264_tables.inc, probed from the compiled reference)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_tables():
    txt = open(os.path.join(ROOT, "h264_video_decoder_demo_b200", "csrc", "host", "h264_tables.inc")).read()

    def arr(name, shape):
        m = re.search(name + r"[^=]*=\s*(\{.*?\});", txt, flags=re.S)
        nums = [int(x) for x in re.findall(r"-?\d+", m.group(1))]
        return np.array(nums).reshape(shape)

    return {"coeff_token": arr("kCoeffToken", (6, 17, 4, 2)), "total_zeros": arr("kTotalZeros", (3, 16, 16, 2)),
            "run_before": arr("kRunBefore", (8, 15, 2)), "me_cbp": arr("kMeCbp", (48, 2))}


T = _load_tables()
BLK_X = [0, 4, 0, 4, 8, 12, 8, 12, 0, 4, 0, 4, 8, 12, 8, 12]
BLK_Y = [0, 0, 4, 4, 0, 0, 4, 4, 8, 8, 12, 12, 8, 8, 12, 12]


class Bits:
    def __init__(self):
        self.b = []

    def u(self, n, v):
        for i in range(n - 1, -1, -1):
            self.b.append((v >> i) & 1)

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def code(self, length, value):
        self.u(int(length), int(value))

    def trailing(self):
        self.b.append(1)
        while len(self.b) % 8:
            self.b.append(0)

    def align_zero(self):
        while len(self.b) % 8:
            self.b.append(0)

    def rbsp(self):
        assert len(self.b) % 8 == 0
        return bytes(int("".join(map(str, self.b[i:i + 8])), 2) for i in range(0, len(self.b), 8))


def nal(nal_ref_idc, nal_type, rbsp, long_start=True):
    out = bytearray(b"\x00\x00\x00\x01" if long_start else b"\x00\x00\x01")
    out.append((nal_ref_idc << 5) | nal_type)
    zeros = 0
    for byte in rbsp:
        if zeros >= 2 and byte <= 3:
            out.append(3)
            zeros = 0
        out.append(byte)
        zeros = zeros + 1 if byte == 0 else 0
    return bytes(out)


class Stream:
    """Random syntax for `n_pics` pictures of wmb x hmb macroblocks."""

    def __init__(self, seed, wmb=8, hmb=6, n_pics=5, t8x8=False, weighted=False, n_refs=3, max_slices=3, pcm=True, poc_type=2, bframes=False, bipred_idc=0,
                 mmco=False, mmco5=False, mmco_set=(1, 2, 3, 4, 6), mmco_mod=True, mmco_idr_lt=True, fn_gaps=False):
        self.rng = np.random.default_rng(seed)
        self.wmb, self.hmb, self.n_pics = wmb, hmb, n_pics
        self.t8x8, self.weighted, self.n_refs, self.max_slices, self.pcm, self.poc_type = t8x8, weighted, n_refs, max_slices, pcm, poc_type
        self.bframes, self.bipred_idc = bframes, bipred_idc
        if bframes:
            self.poc_type = 0
        # mmco: adaptive reference marking (memory_management_control_operation 1-4, 6; 5 with mmco5), long-term reference pictures
        # (IDR long_term_reference_flag, MMCO 3 / 6) and reference list modification with short- and long-term picture numbers
        # (H264SliceHeader.cpp:672, H264RefPicList.cpp:1299-1484, 1736-2136); P-only streams, every picture a reference
        self.mmco, self.mmco5, self.mmco_set, self.mmco_mod, self.mmco_idr_lt = mmco, mmco5, tuple(mmco_set), mmco_mod, mmco_idr_lt
        # fn_gaps: frame_num jumps by more than one between pictures (gaps_in_frame_num_value_allowed_flag = 1); the reference's
        # Decoding_process_for_gaps_in_frame_num is an empty stub (H264RefPicList.cpp:1598): no "non-existing" frames are inserted
        self.fn_gaps = fn_gaps
        self.ops_log = []          # marking / list-modification operations written (coverage of the generated stream)
        self.st, self.lt, self.max_lt, self.prev_ref_fn = [], {}, -1, 0      # short-term frame_nums (decode order), long-term indices in use
        self.out = bytearray()
        self.trace = []            # (picture, mb address, kind) of every macroblock written, for debugging

    # ---------------------------------------------------------------- parameter sets
    def sps(self):
        b = Bits()
        b.u(8, 100); b.u(8, 0); b.u(8, 30); b.ue(0)                 # High profile (8x8 transform allowed), level 3
        b.ue(1); b.ue(0); b.ue(0); b.u(1, 0); b.u(1, 0)             # 4:2:0, 8 bit, no bypass, no scaling matrix
        b.ue(4)                                                      # log2_max_frame_num_minus4 -> 8 bits
        b.ue(self.poc_type)
        if self.poc_type == 0:
            b.ue(4)
        b.ue(self.n_refs); b.u(1, 1 if self.fn_gaps else 0)        # max_num_ref_frames, gaps_in_frame_num_value_allowed_flag
        b.ue(self.wmb - 1); b.ue(self.hmb - 1)
        b.u(1, 1); b.u(1, 1); b.u(1, 0)                              # frame_mbs_only, direct_8x8_inference, no cropping
        b.u(1, 1)                                                    # VUI: only the bitstream restriction (max_num_reorder_frames = 0)
        for _ in range(4):
            b.u(1, 0)
        b.u(1, 0); b.u(1, 0); b.u(1, 0); b.u(1, 0)                   # timing, nal hrd, vcl hrd, pic_struct
        b.u(1, 1); b.u(1, 1); b.ue(0); b.ue(0); b.ue(10); b.ue(10); b.ue(2 if self.bframes else 0); b.ue(self.n_refs)
        b.trailing()
        return nal(3, 7, b.rbsp())

    def pps(self):
        b = Bits()
        b.ue(0); b.ue(0); b.u(1, 0); b.u(1, 0); b.ue(0)              # CAVLC, no bottom-field poc, one slice group
        b.ue(self.n_refs - 1); b.ue(0)
        b.u(1, 1 if self.weighted else 0); b.u(2, self.bipred_idc)
        b.se(0); b.se(0); b.se(int(self.rng.integers(-3, 4)))
        b.u(1, 1); b.u(1, 0); b.u(1, 0)                              # deblocking control present, no constrained intra, no redundant pics
        b.u(1, 1 if self.t8x8 else 0); b.u(1, 0); b.se(int(self.rng.integers(-3, 4)))
        b.trailing()
        return nal(3, 8, b.rbsp())

    # ---------------------------------------------------------------- residual
    def _nC(self, tc_map, avail, mbx, mby, bx, by, slice_of, cur_slice):
        """tc_map: per-4x4 TotalCoeff array of the picture plane; bx,by: block coordinates in the plane."""
        vals = []
        for dx, dy in ((-1, 0), (0, -1)):
            x, y = bx + dx, by + dy
            if x < 0 or y < 0:
                continue
            per = tc_map.shape[1] // self.wmb
            if slice_of[y // per, x // per] != cur_slice:
                continue
            vals.append(int(tc_map[y, x]))
        if len(vals) == 2:
            return (vals[0] + vals[1] + 1) >> 1
        return vals[0] if vals else 0

    def _levels(self, n, maxc):
        """n non-zero levels placed in a list of maxc coefficients."""
        lv = [0] * maxc
        pos = sorted(self.rng.choice(maxc, size=n, replace=False).tolist())
        for p in pos:
            mag = 1 if self.rng.random() < 0.6 else int(self.rng.integers(1, 24))
            lv[p] = mag if self.rng.random() < 0.5 else -mag
        return lv

    def _write_block(self, b, lv, nC, maxc):
        nz = [(i, v) for i, v in enumerate(lv) if v]
        tc = len(nz)
        t1 = 0
        for _, v in reversed(nz):
            if abs(v) == 1 and t1 < 3:
                t1 += 1
            else:
                break
        cls = 4 if nC < 0 else 0 if nC < 2 else 1 if nC < 4 else 2 if nC < 8 else 3
        ln, code = T["coeff_token"][cls][tc][t1]
        assert ln > 0
        b.code(ln, code)
        if tc == 0:
            return 0
        levels = [v for _, v in reversed(nz)]                      # highest frequency first
        for v in levels[:t1]:
            b.u(1, 1 if v < 0 else 0)
        sl = 1 if (tc > 10 and t1 < 3) else 0
        for i, v in enumerate(levels[t1:]):
            code_ = 2 * v - 2 if v > 0 else -2 * v - 1
            if i == 0 and t1 < 3:
                code_ -= 2
            if sl == 0:
                if code_ < 14:
                    b.u(code_, 0); b.u(1, 1)
                elif code_ < 30:
                    b.u(14, 0); b.u(1, 1); b.u(4, code_ - 14)
                else:
                    b.u(15, 0); b.u(1, 1); b.u(12, code_ - 30)
            else:
                if (code_ >> sl) < 15:
                    b.u(code_ >> sl, 0); b.u(1, 1); b.u(sl, code_ & ((1 << sl) - 1))
                else:
                    b.u(15, 0); b.u(1, 1); b.u(12, code_ - (15 << sl))
            if sl == 0:
                sl = 1
            if abs(v) > (3 << (sl - 1)) and sl < 6:
                sl += 1
        last = nz[-1][0]
        total_zeros = last + 1 - tc
        if tc < maxc:
            kind = 1 if maxc == 4 else 0
            ln, code = T["total_zeros"][kind][tc][total_zeros]
            assert ln > 0, (kind, tc, total_zeros)
            b.code(ln, code)
        zeros_left = total_zeros
        idx = [i for i, _ in nz]
        for k in range(tc - 1, 0, -1):
            if zeros_left <= 0:
                break
            run = idx[k] - idx[k - 1] - 1
            ln, code = T["run_before"][min(zeros_left, 7)][run]
            assert ln > 0
            b.code(ln, code)
            zeros_left -= run
        return tc

    # ---------------------------------------------------------------- one picture
    def picture(self, pic_idx, frame_num, kind="P", poc=None, is_ref=True, refs_before=0, refs_after=0):
        """kind: "I" (IDR when pic_idx == 0), "P" or "B"; refs_before / refs_after: reference pictures available with a smaller /
        larger POC (B pictures)."""
        rng = self.rng
        idr = pic_idx == 0
        is_b = kind == "B"
        poc = 2 * pic_idx if poc is None else poc
        n_mbs = self.wmb * self.hmb
        n_slices = int(rng.integers(1, self.max_slices + 1))
        cuts = sorted(set([0] + rng.choice(np.arange(1, n_mbs), size=n_slices - 1, replace=False).tolist())) if n_slices > 1 else [0]
        slice_of = np.zeros((self.hmb, self.wmb), dtype=np.int32)
        for s, c in enumerate(cuts):
            for a in range(c, cuts[s + 1] if s + 1 < len(cuts) else n_mbs):
                slice_of[a // self.wmb, a % self.wmb] = s
        tcY = np.zeros((self.hmb * 4, self.wmb * 4), dtype=np.int32)
        tcC = [np.zeros((self.hmb * 2, self.wmb * 2), dtype=np.int32) for _ in range(2)]
        n_avail_refs = min(refs_before + refs_after, self.n_refs) if self.bframes else min(pic_idx, self.n_refs)
        if is_b and self.bipred_idc == 2:
            n_avail_refs = 1         # implicit weights: the reference divides by the POC distance of the two references before testing it
                                     # for 0 (IP:2957), so both lists keep one entry: the nearest picture before / after
        # The reference builds the reference lists once per picture, from the FIRST slice's header (H264SliceData.cpp:84-124), and
        # its CAVLC te() range for ref_idx comes from that list length (H264MacroBlock.cpp:1321): all slices of a picture share
        # the slice type and num_ref_idx_active here, otherwise the reference itself loses synchronisation.
        pic_is_p = (not idr) and (is_b or kind == "P") and (self.bframes or rng.random() < 0.85)
        pic_n_act = int(rng.integers(1, n_avail_refs + 1)) if pic_is_p else 1
        pic_n_act1 = int(rng.integers(1, n_avail_refs + 1)) if is_b else 1
        mod_ops, mark_ops, idr_lt = None, None, 0
        self._last_was_mmco5 = False
        if self.mmco:
            MAXFN = 256
            if idr:
                self.st, self.lt, self.max_lt = [], {}, -1
                idr_lt = int(rng.integers(0, 2)) if self.mmco_idr_lt else 0
            else:
                n_avail_refs = len(self.st) + len(self.lt)
                pic_is_p = n_avail_refs > 0
                pic_n_act = int(rng.integers(1, n_avail_refs + 1)) if pic_is_p else 1
                picnum = lambda fn: fn if fn <= frame_num else fn - MAXFN
                if pic_is_p and self.mmco_mod and rng.random() < 0.6:    # ref_pic_list_modification_l0
                    mod_ops, pred = [], frame_num
                    for _ in range(int(rng.integers(1, pic_n_act + 1))):
                        if self.lt and (not self.st or rng.random() < 0.4):
                            mod_ops.append((2, int(rng.choice(sorted(self.lt)))))
                        elif self.st:
                            t = int(rng.choice(self.st))
                            tn = picnum(t) % MAXFN                       # picNumL0NoWrap domain
                            delta = tn - pred
                            if delta == 0:
                                continue
                            mod_ops.append((0, -delta - 1) if delta < 0 else (1, delta - 1))
                            pred = tn
                # dec_ref_pic_marking: simulate the marking so that every operation is valid and the DPB never overflows
                cap = self.n_refs
                cur_long = None
                if self.mmco5 and pic_idx >= 2 and rng.random() < 0.35:
                    mark_ops = [(5,)]
                    self._last_was_mmco5 = True
                    self.st, self.lt, self.max_lt = [0], {}, -1           # the picture itself stays, as a short-term reference with frame_num 0
                elif rng.random() < 0.65 or (not self.st and len(self.lt) >= cap):
                    mark_ops = []
                    for _ in range(int(rng.integers(0, 4))):
                        op = int(rng.choice(self.mmco_set))
                        if op == 1 and self.st:
                            t = int(rng.choice(self.st)); mark_ops.append((1, frame_num - picnum(t) - 1)); self.st.remove(t)
                        elif op == 2 and self.lt:
                            i_ = int(rng.choice(sorted(self.lt))); mark_ops.append((2, i_)); del self.lt[i_]
                        elif op == 3 and self.st and self.max_lt >= 0:
                            t = int(rng.choice(self.st)); i_ = int(rng.integers(0, self.max_lt + 1))
                            mark_ops.append((3, frame_num - picnum(t) - 1, i_)); self.st.remove(t); self.lt[i_] = True
                        elif op == 4:
                            v = int(rng.integers(0, cap + 1)); mark_ops.append((4, v))
                            for i_ in [k_ for k_ in self.lt if k_ >= v]:
                                del self.lt[i_]
                            self.max_lt = v - 1
                            if cur_long is not None and cur_long >= v:
                                mark_ops.pop(); self.max_lt = max(self.max_lt, cur_long)       # keep the stream conforming: do not free the current picture
                        elif op == 6 and self.max_lt >= 0 and cur_long is None:
                            i_ = int(rng.integers(0, self.max_lt + 1)); mark_ops.append((6, i_)); self.lt[i_] = True; cur_long = i_
                    while len(self.st) + len(self.lt) + (0 if cur_long is not None else 1) > cap:
                        if self.st:
                            t = self.st[0]; mark_ops.append((1, frame_num - picnum(t) - 1)); self.st.remove(t)
                        else:
                            i_ = [k_ for k_ in sorted(self.lt) if k_ != cur_long][0]; mark_ops.append((2, i_)); del self.lt[i_]
                    if cur_long is None:
                        self.st.append(frame_num)
                else:                                                     # sliding window
                    if len(self.st) + len(self.lt) >= cap:
                        self.st.remove(min(self.st, key=picnum))
                    self.st.append(frame_num)
            self.ops_log += [("mod", op[0]) for op in (mod_ops or [])] + [("mmco", op[0]) for op in (mark_ops or [])] + ([("idr_lt", 1)] if idr and idr_lt else [])
            if idr:
                if idr_lt:
                    self.lt[0] = True; self.max_lt = 0
                else:
                    self.st.append(frame_num)
        for s, first in enumerate(cuts):
            last = (cuts[s + 1] if s + 1 < len(cuts) else n_mbs) - 1
            is_p = pic_is_p
            b = Bits()
            b.ue(first); b.ue((1 if is_b else 0) if is_p else 2); b.ue(0)
            b.u(8, frame_num)
            if idr:
                b.ue(0)
            if self.poc_type == 0:
                b.u(8, poc & 255)
            n_act = n_act1 = 1
            if is_p and is_b:
                n_act, n_act1 = pic_n_act, pic_n_act1
                b.u(1, 1)                                            # direct_spatial_mv_pred_flag
                b.u(1, 1); b.ue(n_act - 1); b.ue(n_act1 - 1)
                b.u(1, 0); b.u(1, 0)                                 # no modification of either list
                if self.bipred_idc == 1:
                    ld, cd = int(rng.integers(0, 6)), int(rng.integers(0, 6))
                    b.ue(ld); b.ue(cd)
                    for cnt in (n_act, n_act1):
                        for _ in range(cnt):
                            if rng.random() < 0.6:
                                b.u(1, 1); b.se(int(rng.integers(-20, 60))); b.se(int(rng.integers(-10, 11)))
                            else:
                                b.u(1, 0)
                            if rng.random() < 0.5:
                                b.u(1, 1)
                                for _ in range(2):
                                    b.se(int(rng.integers(-20, 60))); b.se(int(rng.integers(-10, 11)))
                            else:
                                b.u(1, 0)
            elif is_p:
                n_act = pic_n_act
                b.u(1, 1); b.ue(n_act - 1)                          # num_ref_idx_active_override
                if mod_ops:
                    b.u(1, 1)
                    for op in mod_ops:
                        b.ue(op[0]); b.ue(op[1])
                    b.ue(3)
                else:
                    b.u(1, 0)                                        # no list modification
                if self.weighted:
                    ld, cd = int(rng.integers(0, 6)), int(rng.integers(0, 6))
                    b.ue(ld); b.ue(cd)
                    for _ in range(n_act):
                        if rng.random() < 0.6:
                            b.u(1, 1); b.se(int(rng.integers(-20, 60))); b.se(int(rng.integers(-10, 11)))
                        else:
                            b.u(1, 0)
                        if rng.random() < 0.5:
                            b.u(1, 1)
                            for _ in range(2):
                                b.se(int(rng.integers(-20, 60))); b.se(int(rng.integers(-10, 11)))
                        else:
                            b.u(1, 0)
            if idr:
                b.u(1, 0); b.u(1, idr_lt)                            # no_output_of_prior_pics_flag, long_term_reference_flag
            elif is_ref and mark_ops is not None:
                b.u(1, 1)                                            # adaptive_ref_pic_marking_mode_flag
                for op in mark_ops:
                    b.ue(op[0])
                    for v_ in op[1:]:
                        b.ue(v_)
                b.ue(0)
            elif is_ref:
                b.u(1, 0)                                            # sliding window
            qp = 26 + int(rng.integers(-8, 9))
            b.se(qp - 26)
            idc = int(rng.integers(0, 3))
            b.ue(idc)
            if idc != 1:
                b.se(int(rng.integers(-3, 4))); b.se(int(rng.integers(-3, 4)))
            # ---- slice data
            skip_run = 0
            for a in range(first, last + 1):
                mbx, mby = a % self.wmb, a // self.wmb
                left_ok = mbx > 0 and slice_of[mby, mbx - 1] == s
                top_ok = mby > 0 and slice_of[mby - 1, mbx] == s
                if is_p and rng.random() < 0.25:
                    skip_run += 1
                    self.trace.append((pic_idx, a, "skip"))
                    continue
                if is_p:
                    b.ue(skip_run); skip_run = 0
                kind = rng.random()
                self.trace.append((pic_idx, a, "inter" if (is_p and kind < 0.7) else "intra"))
                kind = float(kind)
                if is_p and is_b and kind < 0.75:
                    qp = self._b_mb(b, n_act, n_act1, qp, tcY, tcC, mbx, mby, slice_of, s)
                elif is_p and kind < 0.7:
                    qp = self._inter_mb(b, n_act, qp, tcY, tcC, mbx, mby, slice_of, s)
                else:
                    qp = self._intra_mb(b, is_p, qp, tcY, tcC, mbx, mby, slice_of, s, left_ok, top_ok, is_b)
            if is_p and skip_run:
                b.ue(skip_run)
            b.trailing()
            self.out += nal(1 if is_ref else 0, 5 if idr else 1, b.rbsp(), long_start=(s == 0))

    def _residual(self, b, cbp_luma, cbp_chroma, i16, t8, qp, tcY, tcC, mbx, mby, slice_of, s):
        rng = self.rng
        if i16:
            n = int(rng.integers(0, 7))
            nC = self._nC(tcY, None, mbx, mby, mbx * 4, mby * 4, slice_of, s)
            self._write_block(b, self._levels(n, 16), nC, 16)
        for i8 in range(4):
            if not (cbp_luma >> i8) & 1:
                continue
            for i4 in range(4):
                blk = i8 * 4 + i4
                bx, by = mbx * 4 + BLK_X[blk] // 4, mby * 4 + BLK_Y[blk] // 4
                nC = self._nC(tcY, None, mbx, mby, bx, by, slice_of, s)
                maxc = 15 if i16 else 16
                n = int(rng.integers(0, 5)) if rng.random() < 0.8 else int(rng.integers(5, maxc + 1))
                tcY[by, bx] = self._write_block(b, self._levels(n, maxc), nC, maxc)
        if cbp_chroma:
            for c in range(2):
                self._write_block(b, self._levels(int(rng.integers(0, 4)), 4), -1, 4)
        if cbp_chroma == 2:
            for c in range(2):
                for blk in range(4):
                    bx, by = mbx * 2 + blk % 2, mby * 2 + blk // 2
                    nC = self._nC(tcC[c], None, mbx, mby, bx, by, slice_of, s)
                    n = int(rng.integers(0, 4))
                    tcC[c][by, bx] = self._write_block(b, self._levels(n, 15), nC, 15)

    def _cbp_and_residual(self, b, intra_nxn, t8_allowed, qp, tcY, tcC, mbx, mby, slice_of, s, t8_known=None):
        rng = self.rng
        cbp_luma = int(rng.integers(0, 16)) if rng.random() < 0.8 else 0
        cbp_chroma = int(rng.integers(0, 3))
        cbp = cbp_luma | (cbp_chroma << 4)
        col = 0 if intra_nxn else 1
        code_num = int(np.nonzero(T["me_cbp"][:, col] == cbp)[0][0])
        b.ue(code_num)
        t8 = t8_known
        if t8_known is None and cbp_luma > 0 and self.t8x8 and t8_allowed:
            t8 = int(rng.random() < 0.5)
            b.u(1, t8)
        if cbp:
            dq = int(rng.integers(-2, 3)) if rng.random() < 0.3 else 0
            if not 12 <= qp + dq <= 44:
                dq = 0
            b.se(dq)
            qp += dq
            self._residual(b, cbp_luma, cbp_chroma, False, t8, qp, tcY, tcC, mbx, mby, slice_of, s)
        return qp

    def _inter_mb(self, b, n_act, qp, tcY, tcC, mbx, mby, slice_of, s):
        rng = self.rng
        t = int(rng.choice([0, 1, 2, 3, 3, 3, 4]))
        b.ue(t)
        no_sub8 = True

        def mvd():
            for _ in range(2):
                b.se(int(rng.integers(-5, 6)) if rng.random() < 0.7 else 0)

        def ref():
            if n_act > 1:
                r = int(rng.integers(0, n_act))
                if n_act == 2:
                    b.u(1, 0 if r else 1)
                else:
                    b.ue(r)
        if t <= 2:
            parts = 1 if t == 0 else 2
            for _ in range(parts):
                ref()
            for _ in range(parts):
                mvd()
        else:
            subs = [int(rng.integers(0, 4)) for _ in range(4)]
            for st in subs:
                b.ue(st)
                if st:
                    no_sub8 = False
            if t == 3:
                for _ in range(4):
                    ref()
            for st in subs:
                for _ in range((1, 2, 2, 4)[st]):
                    mvd()
        return self._cbp_and_residual(b, False, no_sub8, qp, tcY, tcC, mbx, mby, slice_of, s)

    def _b_mb(self, b, n0, n1, qp, tcY, tcC, mbx, mby, slice_of, s):
        """A B macroblock: B_Direct_16x16, 16x16 / 16x8 / 8x16 with every list combination, B_8x8 with direct / L0 / L1 / Bi 8x8
        sub-macroblocks (the reference rejects the smaller B sub-partitions, H264MacroBlock.cpp:1061)."""
        rng = self.rng
        L0, L1, BI = 1, 2, 3
        pairs = [(L0, L0), (L1, L1), (L0, L1), (L1, L0), (L0, BI), (L1, BI), (BI, L0), (BI, L1), (BI, BI)]
        t = int(rng.choice([0, 1, 2, 3] + list(range(4, 22)) + [22, 22, 22]))
        b.ue(t)

        def ref(n):
            if n > 1:
                r = int(rng.integers(0, n))
                if n == 2:
                    b.u(1, 0 if r else 1)
                else:
                    b.ue(r)

        def mvd():
            for _ in range(2):
                b.se(int(rng.integers(-5, 6)) if rng.random() < 0.7 else 0)
        if t == 0:
            parts = []
        elif t <= 3:
            parts = [t]
        elif t < 22:
            parts = list(pairs[(t - 4) // 2])
        else:
            subs = [int(rng.integers(0, 4)) for _ in range(4)]
            for st in subs:
                b.ue(st)
            parts = [st for st in subs if st]                        # sub type 1/2/3 = L0/L1/Bi 8x8; 0 = direct (nothing coded)
        for p in parts:
            if p & L0:
                ref(n0)
        for p in parts:
            if p & L1:
                ref(n1)
        for p in parts:
            if p & L0:
                mvd()
        for p in parts:
            if p & L1:
                mvd()
        return self._cbp_and_residual(b, False, True, qp, tcY, tcC, mbx, mby, slice_of, s)

    def _intra_mb(self, b, is_p, qp, tcY, tcC, mbx, mby, slice_of, s, left_ok, top_ok, is_b=False):
        rng = self.rng
        off = (23 if is_b else 5) if is_p else 0
        kind = rng.random()
        if self.pcm and kind < 0.08:
            b.ue(off + 25)
            b.align_zero()
            for _ in range(384):
                b.u(8, int(rng.integers(0, 256)))
            tcY[mby * 4:mby * 4 + 4, mbx * 4:mbx * 4 + 4] = 0
            return qp
        diag_ok = left_ok and top_ok and slice_of[mby - 1, mbx - 1] == s
        chroma_mode = int(rng.integers(0, 4 if diag_ok else 3)) if (left_ok and top_ok) else 0
        if kind < 0.5:
            modes = [2]
            if top_ok:
                modes.append(0)
            if left_ok:
                modes.append(1)
            if diag_ok:
                modes.append(3)
            pm = int(rng.choice(modes))
            cbp_chroma = int(rng.integers(0, 3))
            cbp_luma = 15 if rng.random() < 0.5 else 0
            b.ue(off + 1 + pm + 4 * cbp_chroma + (12 if cbp_luma else 0))
            b.ue(chroma_mode)
            dq = int(rng.integers(-2, 3)) if rng.random() < 0.3 else 0
            if not 12 <= qp + dq <= 44:
                dq = 0
            b.se(dq)
            qp += dq
            self._residual(b, cbp_luma, cbp_chroma, True, 0, qp, tcY, tcC, mbx, mby, slice_of, s)
            return qp
        b.ue(off + 0)
        t8 = 0
        if self.t8x8:
            t8 = int(rng.random() < 0.5)
            b.u(1, t8)
        for blk in range(4 if t8 else 16):
            inner = (blk == 3) if t8 else (BLK_X[blk] > 0 and BLK_Y[blk] > 0)
            if inner and rng.random() < 0.6:
                b.u(1, 0); b.u(3, int(rng.integers(0, 8)))
            else:
                b.u(1, 1)
        b.ue(chroma_mode)
        return self._cbp_and_residual(b, True, False, qp, tcY, tcC, mbx, mby, slice_of, s, t8_known=t8)

    def build(self):
        self.out += self.sps() + self.pps()
        if self.mmco:
            fn = 0
            for p in range(self.n_pics):
                self.picture(p, fn)
                fn = 1 if self._last_was_mmco5 else (fn + 1) & 255      # after MMCO 5 the picture counts as frame_num 0 (7.4.3)
            return bytes(self.out)
        if not self.bframes:
            fn = 0
            for p in range(self.n_pics):
                self.picture(p, fn & 255)
                fn += 1 + (int(self.rng.integers(1, 4)) if self.fn_gaps and self.rng.random() < 0.5 else 0)
            return bytes(self.out)
        # decode order I0 P4 B2 P8 B6 ... (POC = 2 x display index); every second B picture is itself a reference (B pyramid)
        n_ref_done, ref_pocs, k, display = 0, [], 0, 0
        plan = [("I", 0, True)]
        d = 2
        while len(plan) < self.n_pics:
            plan.append(("P", 2 * d, True))
            if len(plan) < self.n_pics:
                plan.append(("B", 2 * d - 2, bool(self.rng.integers(0, 2))))
            d += 2
        for i, (kind, poc, is_ref) in enumerate(plan):
            window = ref_pocs[-self.n_refs:]
            before, after = sum(1 for q in window if q < poc), sum(1 for q in window if q > poc)
            if kind == "B" and (before + after < 2 or (self.bipred_idc == 2 and (before < 1 or after < 1))):
                kind = "P"
            self.picture(i, n_ref_done & 255, kind=kind, poc=poc, is_ref=is_ref, refs_before=before, refs_after=after)
            if is_ref:
                n_ref_done += 1
                ref_pocs.append(poc)
        return bytes(self.out)


def random_cabac_stream(seed, kind="I", wmb=6, hmb=5, n_pics=4, nbytes=300, t8x8=False):
    """CABAC pictures whose slice DATA is random bytes: valid parameter sets and slice headers, then noise for the arithmetic decoder.
    Whatever syntax the UNMODIFIED reference decodes from the noise (macroblock types incl. I_PCM, sub-partitions, the B sub-types it
    rejects, over-long mb_qp_delta / level / mvd codes, slices that run out of data) is the golden answer — error paths included.
    kind: "I" (intra pictures only), "P" (one reference, ref_idx never coded) or "B" (I P B P B, one reference per list)."""
    b_mode = kind == "B"
    st = Stream(seed=seed, wmb=wmb, hmb=hmb, t8x8=t8x8, n_refs=2 if b_mode else 1, poc_type=0 if b_mode else 2, bframes=b_mode)
    rng = st.rng
    out = bytearray(st.sps())
    b = Bits()
    b.ue(0); b.ue(0); b.u(1, 1); b.u(1, 0); b.ue(0); b.ue(0); b.ue(0)
    b.u(1, 0); b.u(2, int(rng.choice([0, 2])) if b_mode else 0)
    b.se(0); b.se(0); b.se(0); b.u(1, 1); b.u(1, 0); b.u(1, 0)
    b.u(1, 1 if t8x8 else 0); b.u(1, 0); b.se(0)
    b.trailing()
    out += nal(3, 8, b.rbsp())
    if b_mode:
        plan = [("I", 0, True), ("P", 4, True), ("B", 2, False), ("P", 8, True), ("B", 6, False)][:n_pics + 1]
    elif kind == "P":
        plan = [("I", 0, True)] + [("P", 2 * i, True) for i in range(1, n_pics)]
    else:
        plan = [("I", 2 * i, True) for i in range(n_pics)]
    n_ref = 0
    for p, (k, poc, is_ref) in enumerate(plan):
        b = Bits()
        b.ue(0); b.ue({"I": 2, "P": 0, "B": 1}[k]); b.ue(0); b.u(8, n_ref & 255)
        if p == 0:
            b.ue(0)
        if b_mode:
            b.u(8, poc & 255)
        if k == "B":
            b.u(1, 1)                      # direct_spatial_mv_pred_flag
        if k != "I":
            b.u(1, 0); b.u(1, 0)           # no num_ref_idx override, no modification of list 0
        if k == "B":
            b.u(1, 0)
        if p == 0:
            b.u(1, 0); b.u(1, 0)
        elif is_ref:
            b.u(1, 0)
        if k != "I":
            b.ue(int(rng.integers(0, 3)))  # cabac_init_idc
        b.se(int(rng.integers(-6, 7))); b.ue(0); b.se(0); b.se(0)
        while len(b.b) % 8:
            b.b.append(1)                  # cabac_alignment_one_bit
        out += nal(1 if is_ref else 0, 5 if p == 0 else 1, b.rbsp() + rng.integers(0, 256, nbytes, dtype=np.uint8).tobytes() + b"\x80")
        if is_ref:
            n_ref += 1
    return bytes(out)
