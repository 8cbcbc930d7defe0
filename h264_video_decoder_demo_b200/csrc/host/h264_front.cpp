// h264_front.cpp — NAL pump, DPB bookkeeping, output bumping, SoA emission and the C API of the host front end.
// Follows the control flow of the reference: H264VideoDecoder.cpp:48-432 (NAL loop, new-picture detection, EOF handling,
// do_callback), H264PictureBase.cpp:640-712 (end_decode_the_picture_and_get_a_new_empty_picture, getOneEmptyPicture),
// H264PicturesGOP.cpp:89-165 (output bumping), H264RefPicList.cpp (POC, picture numbers, list initialisation and
// modification, marking) — restated, including the places where the reference deviates from H.264 (noted inline).
#include "h264_decoder.h"
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>

namespace h264b2 {

Front::~Front() {
    std::lock_guard<std::mutex> lk(block_mu);
    for (auto &b : free_blocks) { if (free_fn) free_fn(alloc_user, b.p); else free(b.p); }
    for (auto &b : live_blocks) { if (free_fn) free_fn(alloc_user, b.p); else free(b.p); }
}
uint8_t *Front::get_block(size_t bytes, size_t *cap) {
    std::lock_guard<std::mutex> lk(block_mu);
    for (size_t i = 0; i < free_blocks.size(); i++)
        if (free_blocks[i].cap >= bytes) { Block b = free_blocks[i]; free_blocks.erase(free_blocks.begin() + i); live_blocks.push_back(b); *cap = b.cap; return b.p; }
    const size_t want = (bytes + (1u << 20)) & ~(size_t)((1u << 16) - 1);     // headroom: coefficient volume varies per picture
    uint8_t *p = (uint8_t *)(alloc ? alloc(alloc_user, want) : malloc(want));
    if (!p) return nullptr;
    live_blocks.push_back(Block{p, want}); *cap = want;
    return p;
}
void Front::release_block(void *p) {
    std::lock_guard<std::mutex> lk(block_mu);
    for (size_t i = 0; i < live_blocks.size(); i++) if (live_blocks[i].p == p) { free_blocks.push_back(live_blocks[i]); live_blocks.erase(live_blocks.begin() + i); return; }
}

// ------------------------------------------------------------------ NAL pump
static bool find_start(const uint8_t *d, size_t n, size_t from, size_t *pos, int *len) {      // FileReader.cpp:77-124
    for (size_t p = from; p + 3 <= n; p++)
        if (d[p] == 0 && d[p + 1] == 0 && d[p + 2] == 1) {
            if (p >= 1 && p - 1 >= from && d[p - 1] == 0) { *pos = p - 1; *len = 4; } else { *pos = p; *len = 3; }
            return true;
        }
    return false;
}

int Front::pump() {
    if (!primed) {          // a GOP shard: take the parameter sets in front of it first
        primed = 1;
        size_t pos = 0, p1; int l1;
        while (find_start(data, range_begin, pos, &p1, &l1)) {
            const size_t start = p1 + l1; size_t p2; int l2; size_t end = range_begin;
            if (find_start(data, range_begin, start, &p2, &l2)) end = p2;
            pos = end;
            if (end <= start) continue;
            const int t = data[start] & 31;
            if (t != 7 && t != 8) continue;
            rbsp.clear();
            for (size_t i = 1; i < end - start; i++) {
                const uint8_t *nal = data + start; const size_t n = end - start;
                if (i + 2 < n && nal[i] == 0 && nal[i + 1] == 0 && nal[i + 2] == 3) { rbsp.push_back(0); rbsp.push_back(0); i += 2; } else rbsp.push_back(nal[i]);
            }
            if (!rbsp.empty()) rbsp.back() &= 0xFE;
            const size_t nbytes = rbsp.size(); rbsp.resize(nbytes + 16, 0); br.init(rbsp.data(), nbytes);
            if (t == 7) { SPS sp; parse_sps(br, sp); if (sp.valid && sp.sps_id >= 0 && sp.sps_id < 32) { spss[sp.sps_id] = sp; max_num_reorder_frames = sp.max_num_reorder_frames; sps_seen = 1; } }
            else { PPS pp; if (parse_pps(br, pp, spss) != -2 && pp.pps_id >= 0 && pp.pps_id < 256) { ppss[pp.pps_id] = pp; pps_seen = 1; } }
        }
        nal_pos = range_begin;
    }
    // Decode NAL units until at least one event is queued.
    while (events.size() == ev_pos) {
        events.clear(); ev_pos = 0;
        if (eof_done) { H264B2FrontEvent e; memset(&e, 0, sizeof e); e.kind = H264B2_EV_END; events.push_back(e); return 0; }
        size_t p1; int l1;
        if (stop || !find_start(data, size, nal_pos, &p1, &l1)) {
            // end of stream: H264VideoDecoder.cpp:354-374 — the current picture goes to the output process WITHOUT the
            // deblocking/marking of end_decode_the_picture (Q1), then everything pending is flushed.
            if (!stop) {
                if (cur >= 0) { if (pic_active) emit_picture(more_follows ? 1 : 0); do_callback(cur, slots[cur].decode_idx >= 0 ? pic_sh.IdrPicFlag : 0); }
                do_callback(-1, 1);
            }
            eof_done = 1;
            continue;
        }
        size_t p2; int l2;
        const size_t start = p1 + l1;
        size_t end = size;
        if (find_start(data, size, start, &p2, &l2)) end = p2;
        nal_pos = end;
        if (end <= start) continue;
        const uint8_t *nal = data + start; const size_t n = end - start;
        const int nal_ref_idc = (nal[0] >> 5) & 3, nal_unit_type = nal[0] & 31;
        if (nal_unit_type == 14 || nal_unit_type == 20 || nal_unit_type == 21) continue;      // SVC/MVC extensions: not decoded
        // RBSP extraction (H264NalUnit.cpp:199-216) + the reference's "last bit of a buffer reads as 0" (Bitstream.cpp:52-56)
        rbsp.clear(); rbsp.reserve(n + 16);
        for (size_t i = 1; i < n; i++) {
            if (i + 2 < n && nal[i] == 0 && nal[i + 1] == 0 && nal[i + 2] == 3) { rbsp.push_back(0); rbsp.push_back(0); i += 2; }
            else rbsp.push_back(nal[i]);
        }
        if (!rbsp.empty()) rbsp.back() &= 0xFE;
        const size_t nbytes = rbsp.size();
        rbsp.resize(nbytes + 16, 0);
        br.init(rbsp.data(), nbytes);
        switch (nal_unit_type) {
        case 1: case 5: {
            if (nal_unit_type == 5 && (!sps_seen || !pps_seen)) break;
            int r = handle_slice_nal(nal_ref_idc, nal_unit_type);
            if (r < 0) return r;
            break; }
        case 7: { SPS s; parse_sps(br, s); if (s.valid && s.sps_id >= 0 && s.sps_id < 32) { spss[s.sps_id] = s; max_num_reorder_frames = s.max_num_reorder_frames; sps_seen = 1; } break; }
        case 8: { PPS p; int r = parse_pps(br, p, spss); if (r == -2) { error = "slice groups (FMO) are not supported"; return -1; } if (p.pps_id >= 0 && p.pps_id < 256) { ppss[p.pps_id] = p; pps_seen = 1; } break; }
        case 2: case 3: case 4: error = "data partitioning NAL units are not supported"; return -1;
        default: break;
        }
    }
    return 0;
}

int Front::handle_slice_nal(int nal_ref_idc, int nal_unit_type) {
    if (parse_slice_header(br, nal_ref_idc, nal_unit_type, spss, ppss, sh) != 0) { error = "slice_header() failed"; return -1; }    // VD:92 RETURN_IF_FAILED
    if (sh.field_pic_flag) { error = "field pictures (PAFF) are not supported (the reference's PAFF path is undefined, SURVEY Q16)"; return -1; }
    if (sh.sps.ChromaArrayType != 1 || sh.sps.bit_depth_luma_minus8 || sh.sps.bit_depth_chroma_minus8) { error = "only 8-bit 4:2:0 is supported"; return -1; }
    const bool is_new = first_vcl_nal_of_picture(sh, last_sh);
    if (is_new && cur >= 0 && cur_has_ptr) {
        int r = end_decode_and_new_picture();
        if (stop) return 0;
        if (r < 0 || cur < 0) { error = "no free picture in the decoded picture buffer"; return -1; }   // VD:127 picture_current == NULL
    }
    if (cur < 0) { error = "no current picture"; return -1; }
    last_sh = sh;
    // ---- CH264Picture::decode_one_slice (H264Picture.cpp:117-172)
    Slot &s = slots[cur];
    if (wmb && (wmb != sh.sps.PicWidthInMbs || hmb != sh.PicHeightInMbs)) { error = "picture size changes inside a stream are not supported"; return -1; }
    wmb = sh.sps.PicWidthInMbs; hmb = sh.PicHeightInMbs; nmb = wmb * hmb;
    s.p_coded_type = 1; s.coded_type = 1;
    cur_has_ptr = 1;
    pic_sh = sh;
    s.hdr_poc_lsb = sh.pic_order_cnt_lsb; s.hdr_frame_num = sh.frame_num; s.hdr_field_pic = sh.field_pic_flag; s.hdr_mbaff_sps = sh.sps.mb_adaptive_frame_field_flag;
    if (!pic_active) start_picture_storage();
    if (decode_slice() == 0) s.in_use = 1;      // a failing slice is skipped (VD:133 CONTINUE_IF_FAILED)
    return 0;
}

void Front::start_picture_storage() {
    Slot &s = slots[cur];
    mbs.resize(nmb); memset(mbs.data(), 0, sizeof(MbT) * nmb);
    info.resize(nmb); memset(info.data(), 0, sizeof(H264B2MbInfo) * nmb);
    modes.assign(nmb, 0); coff.assign(nmb, 0);
    ncoef = 0;
    weights.clear();
    H264B2Weight d; memset(&d, 0, sizeof d); for (int c = 0; c < 3; c++) { d.w0[c] = 1; d.w1[c] = 1; }
    weights.push_back(d);
    s.motion.resize(nmb); memset(s.motion.data(), 0, sizeof(H264B2MbMotion) * nmb);
    s.col.resize(nmb); memset(s.col.data(), 0, sizeof(ColMb) * nmb);
    s.decode_idx = decode_count++;
    has_inter = 0; pic_active = 1;
    mb_skip_flag = 0; mb_field = 0;          // CH264SliceData::init() at picture reset (SD:29-41)
}

// ------------------------------------------------------------------ end of a picture (H264PictureBase.cpp:665-712)
int Front::end_decode_and_new_picture() {
    const int c = cur;
    Slot &s = slots[c];
    s.p_finished = 1;
    if (pic_sh.nal_ref_idc != 0) {
        mark_reference();
        if (s.mmco5) { const int t = s.PicOrderCnt; s.TopFieldOrderCnt -= t; s.BottomFieldOrderCnt -= t; }
    }
    int e = -1;
    for (int i = 0; i < 16; i++)
        if (i != c && slots[i].p_mark != MARK_SHORT && slots[i].p_mark != MARK_LONG && slots[i].in_use == 0) { e = i; break; }
    if (pic_active) emit_picture(1);
    if (e >= 0) {
        Slot &n = slots[e];
        std::vector<H264B2MbMotion> m; m.swap(n.motion); std::vector<ColMb> cl; cl.swap(n.col);
        const int hl = n.hdr_poc_lsb, hf = n.hdr_frame_num, hp = n.hdr_field_pic, hm = n.hdr_mbaff_sps;
        n = Slot();                                   // CH264Picture::reset + CH264PictureBase::reset (the slice header copy survives a reset)
        n.motion.swap(m); n.col.swap(cl);
        n.hdr_poc_lsb = hl; n.hdr_frame_num = hf; n.hdr_field_pic = hp; n.hdr_mbaff_sps = hm;
        n.in_use = 1;
        n.prev = c;
        n.prev_ref = (s.mark == MARK_SHORT || s.mark == MARK_LONG) ? c : s.prev_ref;
    }
    pic_num_cnt++;
    do_callback(c, pic_sh.IdrPicFlag);
    cur = e; cur_has_ptr = 0;
    return e >= 0 ? 0 : -1;
}

// ------------------------------------------------------------------ output (H264VideoDecoder.cpp:380-432, H264PicturesGOP.cpp:89-165)
int Front::get_one_out(int np, int *out) {
    *out = -1;
    if (max_num_reorder_frames == 0) { *out = np; return 0; }
    int index = -1;
    for (int i = 0; i < out_len; i++) if (i == 0 || slots[out_buf[i]].PicOrderCnt < slots[out_buf[index]].PicOrderCnt) index = i;
    if (np < 0) {
        if (out_len > 0) { *out = out_buf[index]; out_len--; for (int i = index; i < out_len; i++) out_buf[i] = out_buf[i + 1]; }
    } else if (out_len < max_num_reorder_frames) out_buf[out_len++] = np;
    else if (slots[np].PicOrderCnt < slots[out_buf[index]].PicOrderCnt) *out = np;
    else { *out = out_buf[index]; out_buf[index] = np; }
    return 0;
}

void Front::do_callback(int pic, int flush) {
    auto emit = [&](int o) {
        H264B2FrontEvent e; memset(&e, 0, sizeof e);
        e.kind = H264B2_EV_OUTPUT; e.decode_idx = slots[o].decode_idx; e.surface = o; e.width_mbs = wmb; e.height_mbs = hmb;
        e.hdr.poc = slots[o].PicOrderCnt;
        events.push_back(e);
        slots[o].in_use = 0;
    };
    int o;
    if (flush) for (;;) { get_one_out(-1, &o); if (o < 0) break; emit(o); }
    get_one_out(pic, &o);
    if (o >= 0) emit(o);
}

// ------------------------------------------------------------------ POC (H264RefPicList.cpp:16-282)
int Front::decode_poc() {
    Slot &s = slots[cur];
    const SPS &sps = sh.sps;
    if (sps.pic_order_cnt_type == 0) {
        int prevMsb = 0, prevLsb = 0;
        if (!sh.IdrPicFlag) {
            if (s.prev_ref < 0) return -1;
            const Slot &p = slots[s.prev_ref];
            if (p.mmco5) { prevMsb = 0; prevLsb = p.TopFieldOrderCnt; }
            else { prevMsb = p.PicOrderCntMsb; prevLsb = p.hdr_poc_lsb; }
        }
        const int lsb = sh.pic_order_cnt_lsb, maxl = sps.MaxPicOrderCntLsb;
        if (lsb < prevLsb && (prevLsb - lsb) >= maxl / 2) s.PicOrderCntMsb = prevMsb + maxl;
        else if (lsb > prevLsb && (lsb - prevLsb) > maxl / 2) s.PicOrderCntMsb = prevMsb - maxl;
        else s.PicOrderCntMsb = prevMsb;
        s.TopFieldOrderCnt = s.PicOrderCntMsb + lsb;
        s.BottomFieldOrderCnt = s.TopFieldOrderCnt + sh.delta_pic_order_cnt_bottom;
    } else {
        // types 1 and 2 read the previous picture in decoding order (m_picture_previous, a live pointer in the reference)
        const int pv = s.prev;
        if (!sh.IdrPicFlag && pv < 0) return -1;
        int prevOff = 0;
        if (!sh.IdrPicFlag) prevOff = slots[pv].mmco5 ? 0 : slots[pv].FrameNumOffset;
        if (sh.IdrPicFlag) s.FrameNumOffset = 0;
        else if (slots[pv].hdr_frame_num > sh.frame_num) s.FrameNumOffset = prevOff + sps.MaxFrameNum;
        else s.FrameNumOffset = prevOff;
        if (sps.pic_order_cnt_type == 2) {
            int t = 0;
            if (!sh.IdrPicFlag) t = sh.nal_ref_idc == 0 ? 2 * (s.FrameNumOffset + sh.frame_num) - 1 : 2 * (s.FrameNumOffset + sh.frame_num);
            s.TopFieldOrderCnt = s.BottomFieldOrderCnt = t;
        } else {    // 8.2.1.2
            int absFrameNum = sps.num_ref_frames_in_pic_order_cnt_cycle != 0 ? s.FrameNumOffset + sh.frame_num : 0;
            if (sh.nal_ref_idc == 0 && absFrameNum > 0) absFrameNum--;
            int expected = 0;
            if (absFrameNum > 0) {
                const int cyc = (absFrameNum - 1) / sps.num_ref_frames_in_pic_order_cnt_cycle, in = (absFrameNum - 1) % sps.num_ref_frames_in_pic_order_cnt_cycle;
                expected = cyc * sps.ExpectedDeltaPerPicOrderCntCycle;
                for (int i = 0; i <= in; i++) expected += sps.offset_for_ref_frame[i];
            }
            if (sh.nal_ref_idc == 0) expected += sps.offset_for_non_ref_pic;
            s.TopFieldOrderCnt = expected + sh.delta_pic_order_cnt[0];
            s.BottomFieldOrderCnt = s.TopFieldOrderCnt + sps.offset_for_top_to_bottom_field + sh.delta_pic_order_cnt[1];
        }
    }
    s.PicOrderCnt = std::min(s.TopFieldOrderCnt, s.BottomFieldOrderCnt);
    return 0;
}

// ------------------------------------------------------------------ reference picture lists (H264RefPicList.cpp:284-1480)
void Front::picture_numbers() {
    Slot &c = slots[cur];
    c.FrameNum = sh.frame_num;
    for (int i = 0; i < 16; i++) {
        Slot &s = slots[i];
        if (s.mark == MARK_SHORT) s.FrameNumWrap = s.FrameNum > sh.frame_num ? s.FrameNum - sh.sps.MaxFrameNum : s.FrameNum;
    }
    for (int i = 0; i < 16; i++) {
        Slot &s = slots[i];
        if (s.mark == MARK_SHORT) { s.PicNum = s.FrameNumWrap; s.p_PicNum = s.PicNum; }
        if (s.mark == MARK_LONG) { s.LongTermPicNum = s.LongTermFrameIdx; s.p_LongTermPicNum = s.LongTermPicNum; }
    }
}

template <class Less> static void bubble(int *idx, int n, Less swap_if) {        // the reference's stable bubble sorts
    for (int i = 0; i < n - 1; i++) for (int j = 0; j < n - i - 1; j++) if (swap_if(idx[j], idx[j + 1])) std::swap(idx[j], idx[j + 1]);
}

int Front::init_lists() {
    Slot &c = slots[cur];
    const int st = sh.slice_type % 5;
    int lens[2] = {c.listlen[0], c.listlen[1]};
    if (st == SLICE_P || st == SLICE_SP) {
        int sh_[16], lg[16], ns = 0, nl = 0;
        for (int i = 0; i < 16; i++) { if (slots[i].mark == MARK_SHORT) sh_[ns++] = i; else if (slots[i].mark == MARK_LONG) lg[nl++] = i; }
        if (ns + nl == 0) return -1;
        bubble(sh_, ns, [&](int a, int b) { return slots[a].PicNum < slots[b].PicNum; });
        bubble(lg, nl, [&](int a, int b) { return slots[a].LongTermPicNum > slots[b].LongTermPicNum; });
        int j = 0;
        for (int i = 0; i < ns; i++) c.list[0][j++] = sh_[i];
        for (int i = 0; i < nl; i++) c.list[0][j++] = lg[i];
        lens[0] = j;
        // (MBAFF: the reference also writes RefPicList0[16 + 2i (+1)], i.e. into m_RefPicList1, RPL:649-660; never read back for P slices)
    } else if (st == SLICE_B) {
        const int poc = c.PicOrderCnt;
        int total = 0;
        for (int l = 0; l < 2; l++) {
            int a[16], b[16], lg[16], na = 0, nb = 0, nl = 0;
            for (int i = 0; i < 16; i++) {
                if (slots[i].mark == MARK_SHORT) { const bool first = l == 0 ? slots[i].PicOrderCnt < poc : slots[i].PicOrderCnt > poc; if (first) a[na++] = i; else b[nb++] = i; }
                else if (slots[i].mark == MARK_LONG) lg[nl++] = i;
            }
            if (na + nb + nl == 0) return -1;
            total = na + nb + nl;
            if (l == 0) { bubble(a, na, [&](int x, int y) { return slots[x].PicOrderCnt < slots[y].PicOrderCnt; }); bubble(b, nb, [&](int x, int y) { return slots[x].PicOrderCnt > slots[y].PicOrderCnt; }); }
            else { bubble(a, na, [&](int x, int y) { return slots[x].PicOrderCnt > slots[y].PicOrderCnt; }); bubble(b, nb, [&](int x, int y) { return slots[x].PicOrderCnt < slots[y].PicOrderCnt; }); }
            bubble(lg, nl, [&](int x, int y) { return slots[x].LongTermPicNum > slots[y].LongTermPicNum; });
            int j = 0;
            for (int i = 0; i < na; i++) c.list[l][j++] = a[i];
            for (int i = 0; i < nb; i++) c.list[l][j++] = b[i];
            for (int i = 0; i < nl; i++) c.list[l][j++] = lg[i];
            lens[l] = j;
        }
        // RPL:1000-1016: "lists identical -> swap the first two entries of list 1"; with a single reference the flag stays 0 and the swap happens too
        int differ = 0;
        if (total > 1) for (int i = 0; i < 16; i++) if (c.list[1][i] != c.list[0][i]) { differ = 1; break; }
        if (!differ) std::swap(c.list[1][0], c.list[1][1]);
    }
    const int act[2] = {sh.num_ref_idx_l0_active_minus1 + 1, sh.num_ref_idx_l1_active_minus1 + 1};
    for (int l = 0; l < 2; l++) {
        if (lens[l] > act[l]) { for (int i = act[l]; i < lens[l] && i < 34; i++) c.list[l][i] = -1; lens[l] = act[l]; }
        if (lens[l] < act[l]) for (int i = lens[l]; i < act[l]; i++) c.list[l][i] = -1;
    }
    c.listlen[0] = lens[0]; c.listlen[1] = lens[1];
    return 0;
}

int Front::modify_lists() {
    Slot &c = slots[cur];
    for (int l = 0; l < 2; l++) {
        if (l == 1 && sh.slice_type != SLICE_B) break;
        if (!sh.ref_pic_list_modification_flag[l]) continue;
        int refIdx = 0, pred = sh.CurrPicNum;
        int *L = c.list[l];
        const int n = l ? sh.num_ref_idx_l1_active_minus1 : sh.num_ref_idx_l0_active_minus1;
        for (int i = 0; i < sh.modification_count[l]; i++) {
            const int idc = sh.modification_of_pic_nums_idc[l][i];
            if (idc == 0 || idc == 1) {
                const int d = sh.abs_diff_pic_num_minus1[l][i] + 1;
                int nowrap;
                if (idc == 0) nowrap = pred - d < 0 ? pred - d + sh.MaxPicNum : pred - d;
                else nowrap = pred + d >= sh.MaxPicNum ? pred + d - sh.MaxPicNum : pred + d;
                pred = nowrap;
                const int picNum = nowrap > sh.CurrPicNum ? nowrap - sh.MaxPicNum : nowrap;
                for (int k = n + 1; k > refIdx; k--) L[k] = L[k - 1];
                int k = 0;          // RPL:1413-1421: the wanted picture is searched in the (shifted) list itself, not in the DPB
                for (; k < n + 1; k++) if (L[k] >= 0 && slots[L[k]].p_PicNum == picNum && slots[L[k]].p_mark == MARK_SHORT) break;
                L[refIdx++] = L[k];
                int nIdx = refIdx;
                for (k = refIdx; k <= n + 1; k++)
                    if (L[k] >= 0) { const int f = slots[L[k]].p_mark == MARK_SHORT ? slots[L[k]].p_PicNum : sh.MaxPicNum; if (f != picNum) L[nIdx++] = L[k]; }
                L[n + 1] = -1;
            } else if (idc == 2) {
                const int lt = sh.long_term_pic_num[l][i];
                for (int k = n + 1; k > refIdx; k--) L[k] = L[k - 1];
                int k = 0;
                for (; k < n + 1; k++) if (L[k] >= 0 && slots[L[k]].p_LongTermPicNum == lt) break;
                L[refIdx++] = L[k];
                int nIdx = refIdx;
                for (k = refIdx; k <= n + 1; k++) {
                    if (L[k] < 0) continue;
                    const int f = slots[L[k]].p_mark == MARK_LONG ? slots[L[k]].p_LongTermPicNum : 2 * (slots[cur].MaxLongTermFrameIdx + 1);
                    if (f != lt) L[nIdx++] = L[k];
                }
            } else break;
        }
    }
    return 0;
}

int Front::build_ref_lists() {
    picture_numbers();
    if (init_lists() != 0) return -1;
    modify_lists();
    slots[cur].listlen[0] = sh.num_ref_idx_l0_active_minus1 + 1;
    slots[cur].listlen[1] = sh.num_ref_idx_l1_active_minus1 + 1;
    return 0;
}

// ------------------------------------------------------------------ marking (H264RefPicList.cpp:1486-2136)
int Front::mark_reference() {
    Slot &c = slots[cur];
    const SliceHeader &h = pic_sh;
    auto unmark = [&](Slot &s) { s.mark = MARK_UNUSED; if (s.p_coded_type == 1) { s.p_mark = MARK_UNUSED; s.p_coded_marked = 0; } };
    if (h.IdrPicFlag) {
        // RPL:1506-1518: only the PARENT-level marks of the other pictures are cleared; their frame-level marks survive until the
        // slot is recycled (they still take part in list initialisation and in the sliding-window count)
        for (int i = 0; i < 16; i++) { slots[i].p_mark = MARK_UNUSED; slots[i].p_coded_marked = 0; if (i != cur) slots[i].p_coded_type = 0; }
        if (!h.long_term_reference_flag) { c.mark = MARK_SHORT; c.MaxLongTermFrameIdx = -1; c.p_mark = MARK_SHORT; }
        else { c.mark = MARK_LONG; c.LongTermFrameIdx = 0; c.MaxLongTermFrameIdx = 0; c.p_mark = MARK_LONG; }
        c.p_coded_marked = 1;
        return 0;
    }
    if (!h.adaptive_ref_pic_marking_mode_flag) {
        int ns = 0, nl = 0;
        for (int i = 0; i < 16; i++) { if (slots[i].mark == MARK_SHORT) ns++; if (slots[i].mark == MARK_LONG) nl++; }
        if (ns + nl == std::max(h.sps.max_num_ref_frames, 1) && ns > 0) {       // RPL:1654: equality, not >=
            int best = -1;
            for (int i = 0; i < 16; i++) if (slots[i].mark == MARK_SHORT && (best < 0 || slots[i].FrameNumWrap < slots[best].FrameNumWrap)) best = i;
            if (best >= 0) unmark(slots[best]);
        }
    } else {
        // 8.2.5.4 as the reference implements it for frames (RPL:1736-2136)
        for (int i = 0; i < h.mmco_count; i++) {
            const Mmco &m = h.mmco[i];
            if (m.op == 0) break;
            if (m.op == 1) {
                const int picNumX = h.CurrPicNum - (m.difference_of_pic_nums_minus1 + 1);
                for (int k = 0; k < 16; k++) if (slots[k].mark == MARK_SHORT && slots[k].PicNum == picNumX) { unmark(slots[k]); break; }
            } else if (m.op == 2) {
                for (int k = 0; k < 16; k++) if (slots[k].mark == MARK_LONG && slots[k].LongTermPicNum == m.long_term_pic_num) { unmark(slots[k]); break; }
            } else if (m.op == 3) {
                const int picNumX = h.CurrPicNum - (m.difference_of_pic_nums_minus1 + 1);
                for (int k = 0; k < 16; k++) if (slots[k].mark == MARK_LONG && slots[k].LongTermFrameIdx == m.long_term_frame_idx) unmark(slots[k]);
                for (int k = 0; k < 16; k++) if (slots[k].mark == MARK_SHORT && slots[k].PicNum == picNumX) { slots[k].mark = MARK_LONG; slots[k].p_mark = MARK_LONG; slots[k].LongTermFrameIdx = m.long_term_frame_idx; break; }
            } else if (m.op == 4) {
                const int maxidx = m.max_long_term_frame_idx_plus1 - 1;
                for (int k = 0; k < 16; k++) if (slots[k].mark == MARK_LONG && slots[k].LongTermFrameIdx > maxidx) unmark(slots[k]);
                c.MaxLongTermFrameIdx = maxidx;
            } else if (m.op == 5) {
                // RPL:2037-2045: EVERY entry of the DPB loses its frame-level and parent-level marks and its marked coded type, whatever it
                // held (the current picture gets its short-term mark back below)
                for (int k = 0; k < 16; k++) { slots[k].mark = MARK_UNUSED; slots[k].p_mark = MARK_UNUSED; slots[k].p_coded_marked = 0; }
                c.MaxLongTermFrameIdx = -1; c.mmco5 = 1;
            } else if (m.op == 6) {
                // RPL:2062: the reference looks for the frame that already holds this index with max_long_term_frame_idx_plus1 - 1 of the
                // SAME entry, which an operation 6 never carries (the entry is zeroed, so it compares with -1): nothing is ever released here
                // RPL:2097-2099: for a frame only the FRAME-level mark becomes long-term; the parent-level mark and its coded type are set
                // in the field branches alone, so the parent keeps what it had (that is what the slot choice and the IDR wipe look at)
                c.mark = MARK_LONG; c.LongTermFrameIdx = m.long_term_frame_idx; c.mmco6 = 1;
            }
        }
    }
    if (!c.mmco6) { c.mark = MARK_SHORT; c.MaxLongTermFrameIdx = -1; c.p_mark = MARK_SHORT; c.p_coded_marked = 1; }
    return 0;
}

// ------------------------------------------------------------------ SoA emission (layout = one record of a picture container)
static const int kNorm4[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};      // Table 8-? normAdjust4x4
static const int kNorm8[6][6] = {{20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31}, {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};

void Front::emit_picture(int deblock_enable) {
    Slot &s = slots[cur];
    const SliceHeader &h = pic_sh;
    int n_na = 0, first_na = nmb;
    for (int a = 0; a < nmb; a++) if (info[a].mb_class == H264B2_MB_NA) { n_na++; if (a < first_na) first_na = a; }
    bool flat = true;
    for (int l = 0; l < 6 && flat; l++) { for (int k = 0; k < 16; k++) if (h.ScalingList4x4[l][k] != 16) flat = false; for (int k = 0; k < 64; k++) if (h.ScalingList8x8[l][k] != 16) flat = false; }
    { uint32_t next = (uint32_t)ncoef; for (int a = nmb - 1; a >= 0; a--) { if (info[a].mb_class == H264B2_MB_NA) coff[a] = next; else next = coff[a]; } }
    if (coefs.size() < ncoef + 4) coefs.resize(ncoef + 4);
    while (ncoef % 4) coefs[ncoef++] = 0;
    const size_t b_info = (size_t)nmb * sizeof(H264B2MbInfo), b_modes = (size_t)nmb * 8, b_coff = (size_t)nmb * 4, b_mot = !has_inter ? 0 : packed_motion ? h264b2_pack_coefs_bound((uint32_t)nmb * (uint32_t)(sizeof(H264B2MbMotion) / 2)) : (size_t)nmb * sizeof(H264B2MbMotion);
    const bool pack = packed_coefs && ncoef != 0;
    size_t pack_slack = 0;
    const size_t b_w = weights.size() * sizeof(H264B2Weight), b_c = pack ? h264b2_pack_coefs_bound((uint32_t)ncoef) : ncoef * 2, b_ls = flat ? 0 : (size_t)(2 * 2 * 6 * 16 + 2 * 2 * 6 * 64) * 2;
    // every array starts on a 64-byte boundary inside the block: the engine keeps host alignment on the device and its kernels
    // read records with 16-byte loads
    auto al = [](size_t n) { return (n + 63) & ~(size_t)63; };
    const size_t total = al(b_info) + al(b_modes) + al(b_coff) + al(b_mot) + al(b_w) + al(b_c) + al(b_ls);
    size_t cap = 0;
    uint8_t *blk = get_block(total, &cap);
    H264B2FrontEvent e; memset(&e, 0, sizeof e);
    e.kind = H264B2_EV_PICTURE; e.decode_idx = s.decode_idx; e.surface = cur; e.width_mbs = wmb; e.height_mbs = hmb;
    H264B2FrontPicHdr &ph = e.hdr;
    ph.decode_idx = s.decode_idx; ph.dst_surface = cur; ph.clear_surface = n_na > 0; ph.has_inter = has_inter; ph.deblock_enable = deblock_enable;
    ph.deblock_stop_mb = first_na; ph.mbaff = h.MbaffFrameFlag; ph.cqp0 = h.pps.chroma_qp_index_offset; ph.cqp1 = h.pps.second_chroma_qp_index_offset;
    ph.n_weights = (int)weights.size(); ph.custom_scaling = !flat; ph.slice_type = h.slice_type; ph.poc = s.PicOrderCnt; ph.n_na = n_na;
    ph.n_coefs = (uint32_t)ncoef; ph.nal_ref_idc = (uint32_t)h.nal_ref_idc;
    H264B2PicParams &p = e.params;
    p.width_mbs = wmb; p.height_mbs = hmb; p.mbaff_frame_flag = ph.mbaff; p.chroma_qp_offset[0] = ph.cqp0; p.chroma_qp_offset[1] = ph.cqp1;
    p.dst_surface = cur; p.clear_surface = ph.clear_surface; p.has_inter = has_inter; p.deblock_enable = deblock_enable; p.deblock_stop_mb = first_na;
    p.n_weights = ph.n_weights; p.n_coefs = ph.n_coefs; p.custom_scaling = ph.custom_scaling;
    if (blk) {
        uint8_t *q = blk;
        memcpy(q, info.data(), b_info); p.mb_info = (const H264B2MbInfo *)q; q += al(b_info);
        memcpy(q, modes.data(), b_modes); p.intra_modes = (const uint64_t *)q; q += al(b_modes);
        memcpy(q, coff.data(), b_coff); p.coef_offset = (const uint32_t *)q; q += al(b_coff);
        if (has_inter && packed_motion) {
            size_t used = 0;
            if (h264b2_pack_motion(s.motion.data(), (uint32_t)nmb, q, b_mot, &used)) error = "motion packing failed";
            p.packed |= H264B2_PACKED_MOTION; pack_slack += al(b_mot) - al(used);
            p.motion = (const H264B2MbMotion *)q; q += al(b_mot);
        } else if (has_inter) { memcpy(q, s.motion.data(), b_mot); p.motion = (const H264B2MbMotion *)q; q += al(b_mot); }
        memcpy(q, weights.data(), b_w); p.weights = (const H264B2Weight *)q; q += al(b_w);
        if (pack) {
            size_t used = 0;
            if (h264b2_pack_coefs(coefs.data(), (uint32_t)ncoef, q, b_c, &used)) error = "coefficient packing failed";
            p.packed |= H264B2_PACKED_COEFS;
            pack_slack += al(b_c) - al(used);
        } else if (b_c) memcpy(q, coefs.data(), b_c);
        p.coefs = (const int16_t *)q; q += al(b_c);
        if (!flat) {
            // LevelScale in LIST order (PB:4852-4989 per scan position; chroma uses the luma list, Q7): [intra/inter][frame/field scan][qP%6][k]
            int16_t *ls4 = (int16_t *)q, *ls8 = ls4 + 2 * 2 * 6 * 16;
            for (int inter = 0; inter < 2; inter++) for (int fld = 0; fld < 2; fld++) for (int m = 0; m < 6; m++) {
                const uint8_t *sc = scan4x4_table(fld);
                for (int k = 0; k < 16; k++) { const int i = sc[k] >> 2, j = sc[k] & 3;
                    const int nrm = (i % 2 == 0 && j % 2 == 0) ? kNorm4[m][0] : (i % 2 == 1 && j % 2 == 1) ? kNorm4[m][1] : kNorm4[m][2];
                    *ls4++ = (int16_t)(h.ScalingList4x4[inter ? 3 : 0][k] * nrm); }
            }
            for (int inter = 0; inter < 2; inter++) for (int fld = 0; fld < 2; fld++) for (int m = 0; m < 6; m++) {
                const uint8_t *sc = scan8x8_table(fld);
                for (int k = 0; k < 64; k++) { const int i = sc[k] >> 3, j = sc[k] & 7; int nrm;
                    if (i % 4 == 0 && j % 4 == 0) nrm = kNorm8[m][0]; else if (i % 2 == 1 && j % 2 == 1) nrm = kNorm8[m][1]; else if (i % 4 == 2 && j % 4 == 2) nrm = kNorm8[m][2];
                    else if ((i % 4 == 0 && j % 2 == 1) || (i % 2 == 1 && j % 4 == 0)) nrm = kNorm8[m][3]; else if ((i % 4 == 0 && j % 4 == 2) || (i % 4 == 2 && j % 4 == 0)) nrm = kNorm8[m][4]; else nrm = kNorm8[m][5];
                    *ls8++ = (int16_t)(h.ScalingList8x8[inter ? 1 : 0][k] * nrm); }
            }
            p.level_scale4 = (const int16_t *)q; p.level_scale8 = (const int16_t *)q + 2 * 2 * 6 * 16;
        }
    }
    e.block = blk; e.block_bytes = total - pack_slack;      // bytes that travel
    // what later pictures read from this one (co-located macroblocks)
    for (int a = 0; a < nmb; a++) {
        const MbT &m = mbs[a]; ColMb &c = s.col[a];
        c.type = m.type; c.intra = m.intra; c.field = m.field; c.pf0 = (uint8_t)(m.pf[0][0] | (m.pf[0][1] << 1) | (m.pf[0][2] << 2) | (m.pf[0][3] << 3));
        memcpy(c.ref, m.ref, sizeof c.ref);
    }
    events.push_back(e);
    pic_active = 0;
}

int Front::next_event(H264B2FrontEvent *ev) {
    if (events.size() == ev_pos) { int r = pump(); if (r < 0) return r; }
    *ev = events[ev_pos++];
    return 0;
}

}  // namespace h264b2

// ------------------------------------------------------------------ C API
using namespace h264b2;
struct H264B2Front { Front f; };

extern "C" int h264b2_front_create(H264B2Front **f, h264b2_front_alloc_fn alloc, h264b2_front_free_fn free_fn, void *user) {
    if (!f) return -1;
    H264B2Front *x = new H264B2Front();
    x->f.alloc = alloc; x->f.free_fn = free_fn; x->f.alloc_user = user;
    *f = x;
    return 0;
}
extern "C" int h264b2_front_destroy(H264B2Front *f) { delete f; return 0; }
extern "C" int h264b2_front_set_packed(H264B2Front *f, int flags) { if (!f) return -1; f->f.packed_coefs = (flags & H264B2_PACKED_COEFS) != 0; f->f.packed_motion = (flags & H264B2_PACKED_MOTION) != 0; return 0; }
extern "C" int h264b2_front_open_memory(H264B2Front *f, const uint8_t *data, size_t bytes) {
    if (!f || !data) return -1;
    f->f.data = data; f->f.size = bytes; f->f.nal_pos = 0;
    return 0;
}
extern "C" int h264b2_front_gop_offsets(const uint8_t *data, size_t bytes, size_t *offsets, int max_gops) {
    if (!data) return -1;
    int n = 0;
    size_t pos = 0, p1, group = 0; int l1; bool in_group = false;
    while (find_start(data, bytes, pos, &p1, &l1)) {
        const size_t start = p1 + l1;
        pos = start;
        if (start >= bytes) break;
        const int t = data[start] & 31;
        if (t >= 6 && t <= 9) { if (!in_group) { group = p1; in_group = true; } continue; }       // SEI / SPS / PPS / AUD in front of a picture
        if (t == 5 && start + 1 < bytes && (data[start + 1] & 0x80)) {                                // IDR slice with first_mb_in_slice == 0
            const size_t off = n == 0 ? 0 : (in_group ? group : p1);
            if (offsets && n < max_gops) offsets[n] = off;
            n++;
        }
        in_group = false;
    }
    return n;
}
extern "C" int h264b2_front_open_range(H264B2Front *f, const uint8_t *data, size_t bytes, size_t begin, size_t end, int more_follows) {
    if (!f || !data || begin > end || end > bytes) return -1;
    f->f.data = data; f->f.size = end; f->f.nal_pos = begin; f->f.range_begin = begin; f->f.more_follows = more_follows; f->f.primed = begin == 0;
    return 0;
}
extern "C" int h264b2_front_open_file(H264B2Front *f, const char *path) {
    if (!f || !path) return -1;
    FILE *fp = fopen(path, "rb");
    if (!fp) { f->f.error = std::string("cannot open ") + path; return -1; }
    fseek(fp, 0, SEEK_END); long n = ftell(fp); fseek(fp, 0, SEEK_SET);
    f->f.file.resize((size_t)n + 16, 0);
    if (n > 0 && fread(f->f.file.data(), 1, (size_t)n, fp) != (size_t)n) { fclose(fp); f->f.error = "short read"; return -1; }
    fclose(fp);
    return h264b2_front_open_memory(f, f->f.file.data(), (size_t)n);
}
extern "C" int h264b2_front_next(H264B2Front *f, H264B2FrontEvent *ev) { if (!f || !ev) return -1; return f->f.next_event(ev); }
extern "C" int h264b2_front_release(H264B2Front *f, void *block) { if (!f) return -1; if (block) f->f.release_block(block); return 0; }
extern "C" int h264b2_front_stream_info(H264B2Front *f, H264B2StreamInfo *info) {
    if (!f || !info) return -1;
    const h264b2::SPS &sps = f->f.pic_sh.sps; const h264b2::PPS &pps = f->f.pic_sh.pps;
    if (!sps.valid) return -1;
    info->profile_idc = sps.profile_idc; info->level_idc = sps.level_idc; info->entropy_coding_mode_flag = pps.entropy_coding_mode_flag;
    info->frame_mbs_only_flag = sps.frame_mbs_only_flag; info->mb_adaptive_frame_field_flag = sps.mb_adaptive_frame_field_flag;
    info->width_mbs = sps.PicWidthInMbs; info->height_mbs = sps.FrameHeightInMbs; info->max_num_ref_frames = sps.max_num_ref_frames;
    info->transform_8x8_mode_flag = pps.transform_8x8_mode_flag;
    info->fps = 25.0;
    if (sps.timing_info_present_flag && sps.num_units_in_tick) info->fps = 1.0 * sps.time_scale / sps.num_units_in_tick / 2.0;      // H264SPS.cpp:346-356
    return 0;
}
extern "C" const char *h264b2_front_last_error(H264B2Front *f) { return f ? f->f.error.c_str() : "null front end"; }

extern "C" int h264b2_front_write_container(const char *h264_path, const char *container_path, int max_pictures) {
    return h264b2_front_write_container_range(h264_path, container_path, max_pictures, 0, 0, 0);
}
extern "C" int h264b2_front_write_container_range(const char *h264_path, const char *container_path, int max_pictures, size_t begin, size_t end, int more_follows) {
    struct FileHdr { char magic[8]; uint32_t version, width_mbs, height_mbs, n_pics, n_out, hdr_bytes, pichdr_bytes, reserved; } fh;
    struct OutRec { int32_t decode_idx, pad; uint64_t sum; };
    H264B2Front *f = nullptr;
    if (h264b2_front_create(&f, nullptr, nullptr, nullptr)) return -1;
    if (h264b2_front_open_file(f, h264_path)) { fprintf(stderr, "%s\n", h264b2_front_last_error(f)); h264b2_front_destroy(f); return -1; }
    if (end > begin && h264b2_front_open_range(f, f->f.data, f->f.size, begin, end, more_follows)) { h264b2_front_destroy(f); return -1; }
    FILE *fo = fopen(container_path, "wb");
    if (!fo) { h264b2_front_destroy(f); return -1; }
    memset(&fh, 0, sizeof fh);
    fwrite(&fh, sizeof fh, 1, fo);
    std::vector<OutRec> outs;
    int n_pics = 0, ret = 0, wmb = 0, hmb = 0;
    for (;;) {
        H264B2FrontEvent ev;
        int r = h264b2_front_next(f, &ev);
        if (r < 0) { fprintf(stderr, "front end error: %s\n", h264b2_front_last_error(f)); ret = r; break; }
        if (ev.kind == H264B2_EV_END) break;
        if (ev.kind == H264B2_EV_PICTURE) {
            wmb = ev.width_mbs; hmb = ev.height_mbs;
            if (max_pictures <= 0 || n_pics < max_pictures) {
                const size_t nmb = (size_t)wmb * hmb; const H264B2PicParams &p = ev.params;
                fwrite(&ev.hdr, sizeof ev.hdr, 1, fo);
                fwrite(p.mb_info, sizeof(H264B2MbInfo), nmb, fo); fwrite(p.intra_modes, 8, nmb, fo); fwrite(p.coef_offset, 4, nmb, fo);
                if (ev.hdr.has_inter) fwrite(p.motion, sizeof(H264B2MbMotion), nmb, fo);
                fwrite(p.weights, sizeof(H264B2Weight), (size_t)ev.hdr.n_weights, fo);
                if (ev.hdr.n_coefs) fwrite(p.coefs, 2, ev.hdr.n_coefs, fo);
                if (ev.hdr.custom_scaling) { fwrite(p.level_scale4, 2, 2 * 2 * 6 * 16, fo); fwrite(p.level_scale8, 2, 2 * 2 * 6 * 64, fo); }
                n_pics++;
            }
            h264b2_front_release(f, ev.block);
        } else if (ev.kind == H264B2_EV_OUTPUT) {
            if (max_pictures <= 0 || ev.decode_idx < max_pictures) { OutRec o; o.decode_idx = ev.decode_idx; o.pad = 0; o.sum = 0; outs.push_back(o); }
        }
    }
    if (!outs.empty()) fwrite(outs.data(), sizeof(OutRec), outs.size(), fo);
    memcpy(fh.magic, "H264B2RP", 8); fh.version = 1; fh.width_mbs = (uint32_t)wmb; fh.height_mbs = (uint32_t)hmb; fh.n_pics = (uint32_t)n_pics; fh.n_out = (uint32_t)outs.size();
    fh.hdr_bytes = sizeof fh; fh.pichdr_bytes = sizeof(H264B2FrontPicHdr);
    fseek(fo, 0, SEEK_SET); fwrite(&fh, sizeof fh, 1, fo); fclose(fo);
    h264b2_front_destroy(f);
    return ret;
}
