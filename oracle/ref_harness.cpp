// ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// Drives the UNMODIFIED reference decoder (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/) through its public API
// (CH264VideoDecoder::open + output callback, H264VideoDecoder.h:22-43) and
//   --yuv F     writes the raw I420 output frames in callback order (golden YUV),
//   --replay F  writes, for every picture in DECODING order, the per-picture
//               structure-of-arrays that the B200 engine consumes
//               (include/h264_recon_b200.h) plus checksums of the reference's
//               pre- and post-deblocking picture; this is what pins both the
//               C restatement (oracle/recon_oracle.c) and the CUDA path,
//   (neither)   just decodes and reports frames/s (the CPU baseline).
//
// The reference has no hook between reconstruction and deblocking, so the
// harness interposes two of its functions at LINK time (ld --wrap, sources
// untouched): CH264PictureBase::Deblocking_filter_process (called once per
// picture from H264PictureBase.cpp:707) and CH264Picture::decode_one_slice
// (H264VideoDecoder.cpp:135).  All "final value" derivations (reference picture
// selection, prediction weights) are obtained by calling the reference's own
// public member functions, so the dump is the reference's state, not a
// re-derivation.
#include "H264VideoDecoder.h"
#include "../include/h264_recon_b200.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include <string>

// ---------------------------------------------------------------- checksum
static uint64_t checksum_bytes(const uint8_t *p, size_t n, uint64_t acc, size_t word0) {
    const uint64_t K = 0x9E3779B97F4A7C15ULL;
    size_t nw = n / 4;
    for (size_t i = 0; i < nw; i++) {
        uint32_t w; memcpy(&w, p + 4 * i, 4);
        acc += ((uint64_t)w + 1) * ((2 * (uint64_t)(word0 + i) + 1) * K);
    }
    return acc;
}
static uint64_t checksum_pic(CH264PictureBase &f) {
    size_t ny = (size_t)f.PicWidthInSamplesL * f.PicHeightInSamplesL;
    size_t nc = (size_t)f.PicWidthInSamplesC * f.PicHeightInSamplesC;
    uint64_t h = checksum_bytes(f.m_pic_buff_luma, ny, 0, 0);
    h = checksum_bytes(f.m_pic_buff_cb, nc, h, ny / 4);
    h = checksum_bytes(f.m_pic_buff_cr, nc, h, (ny + nc) / 4);
    return h;
}

// ---------------------------------------------------------------- state
static FILE *g_yuv = nullptr, *g_replay = nullptr, *g_pixdump = nullptr, *g_bgr = nullptr;
static int g_maxframes = 1 << 30, g_nframes = 0, g_quiet = 0;
static int g_decode_idx = 0;
static CH264Picture *g_cur_pic = nullptr;          // picture that received the last slice
static CH264Picture *g_last_dumped = nullptr;      // picture last seen by the deblock wrap
static int g_last_dumped_numcnt = -1;
static std::map<CH264Picture *, int> g_pic2idx;    // DPB picture -> decode index of its current content
struct OutRec { int32_t decode_idx; int32_t pad; uint64_t sum; };
static std::vector<OutRec> g_out;
static uint64_t g_stream_hash = 0;
static int g_wmb = 0, g_hmb = 0;
static long g_dump_pix_idx = -1;
// every slice's header, in the order slice_data() numbered them: the picture object only keeps the LAST slice's header, but
// weights / slice type are per slice (multi-slice pictures with explicit weighted prediction)
struct WHdr {   // the per-slice fields the weight derivation reads (plain copies: CH264SliceHeader owns malloc'ed maps)
    int32_t slice_type, luma_log2_weight_denom, chroma_log2_weight_denom, luma_weight_l0_flag, luma_weight_l1_flag;
    int32_t luma_weight_l0[32], luma_offset_l0[32], chroma_weight_l0[32][2], chroma_offset_l0[32][2];
    int32_t luma_weight_l1[32], luma_offset_l1[32], chroma_weight_l1[32][2], chroma_offset_l1[32][2];
    void load(const CH264SliceHeader &h) {
        slice_type = h.slice_type; luma_log2_weight_denom = h.luma_log2_weight_denom; chroma_log2_weight_denom = h.chroma_log2_weight_denom;
        luma_weight_l0_flag = h.luma_weight_l0_flag; luma_weight_l1_flag = h.luma_weight_l1_flag;      // read as luma_weight_lX[-1] (Q8)
        memcpy(luma_weight_l0, h.luma_weight_l0, sizeof luma_weight_l0); memcpy(luma_offset_l0, h.luma_offset_l0, sizeof luma_offset_l0);
        memcpy(chroma_weight_l0, h.chroma_weight_l0, sizeof chroma_weight_l0); memcpy(chroma_offset_l0, h.chroma_offset_l0, sizeof chroma_offset_l0);
        memcpy(luma_weight_l1, h.luma_weight_l1, sizeof luma_weight_l1); memcpy(luma_offset_l1, h.luma_offset_l1, sizeof luma_offset_l1);
        memcpy(chroma_weight_l1, h.chroma_weight_l1, sizeof chroma_weight_l1); memcpy(chroma_offset_l1, h.chroma_offset_l1, sizeof chroma_offset_l1);
    }
    void store(CH264SliceHeader &h) const {
        h.slice_type = slice_type; h.luma_log2_weight_denom = luma_log2_weight_denom; h.chroma_log2_weight_denom = chroma_log2_weight_denom;
        h.luma_weight_l0_flag = luma_weight_l0_flag; h.luma_weight_l1_flag = luma_weight_l1_flag;
        memcpy(h.luma_weight_l0, luma_weight_l0, sizeof luma_weight_l0); memcpy(h.luma_offset_l0, luma_offset_l0, sizeof luma_offset_l0);
        memcpy(h.chroma_weight_l0, chroma_weight_l0, sizeof chroma_weight_l0); memcpy(h.chroma_offset_l0, chroma_offset_l0, sizeof chroma_offset_l0);
        memcpy(h.luma_weight_l1, luma_weight_l1, sizeof luma_weight_l1); memcpy(h.luma_offset_l1, luma_offset_l1, sizeof luma_offset_l1);
        memcpy(h.chroma_weight_l1, chroma_weight_l1, sizeof chroma_weight_l1); memcpy(h.chroma_offset_l1, chroma_offset_l1, sizeof chroma_offset_l1);
    }
};
static std::map<CH264Picture *, std::vector<WHdr> > g_slice_hdrs;

struct PicHdr {
    int32_t decode_idx, dst_surface, clear_surface, has_inter, deblock_enable, deblock_stop_mb;
    int32_t mbaff, cqp0, cqp1, n_weights, custom_scaling, slice_type, poc, n_na;
    uint32_t n_coefs, nal_ref_idc;
    uint64_t sum_pre, sum_post;
};

struct WKey { int16_t v[16]; bool operator<(const WKey &o) const { return memcmp(v, o.v, sizeof v) < 0; } };

static int slot_of(CH264PictureBase *pic, CH264Picture *p) {
    if (!p) return -1;
    for (int i = 0; i < 16; i++) if (pic->m_dpb[i] == p) return i;
    return -2;
}

static const int norm4[6][3] = {{10,16,13},{11,18,14},{13,20,16},{14,23,18},{16,25,20},{18,29,23}};
static const int norm8[6][6] = {{20,18,32,19,25,24},{22,19,35,21,28,26},{26,23,42,24,33,31},
                                {28,25,45,26,35,33},{32,28,51,30,40,38},{36,32,58,34,46,43}};

static void put16(std::vector<int16_t> &v, const int32_t *src, int n, int shift_in) {
    // shift_in=1: output[0]=0, output[k]=src[k-1] (AC lists)
    for (int k = 0; k < n; k++) {
        int32_t x = shift_in ? (k == 0 ? 0 : src[k - 1]) : src[k];
        if (x < -32768 || x > 32767) { fprintf(stderr, "FATAL level %d out of int16\n", x); exit(3); }
        v.push_back((int16_t)x);
    }
}
static bool anynz(const int32_t *s, int n) { for (int i = 0; i < n; i++) if (s[i]) return true; return false; }

static int g_cur_hdr_slice = -1;
static WHdr g_saved_hdr;
static void dump_picture(CH264PictureBase *pic, int deblock_enable, uint64_t sum_pre, uint64_t sum_post) {
    CH264SliceHeader &sh = pic->m_h264_slice_header;
    const int W = pic->PicWidthInMbs, nmb = pic->PicSizeInMbs;
    const int mbaff = sh.MbaffFrameFlag;
    g_wmb = W; g_hmb = pic->PicHeightInMbs;
    if (sh.field_pic_flag) { fprintf(stderr, "FATAL field pictures unsupported in replay dump\n"); exit(3); }

    std::vector<H264B2MbInfo> info(nmb);
    std::vector<uint64_t> modes(nmb, 0);
    std::vector<uint32_t> coff(nmb, 0);
    std::vector<H264B2MbMotion> mot(nmb);
    std::vector<int16_t> coefs;
    std::vector<H264B2Weight> weights;
    std::map<WKey, int> wmap;
    memset(info.data(), 0, sizeof(H264B2MbInfo) * nmb);
    memset(mot.data(), 0, sizeof(H264B2MbMotion) * nmb);
    { H264B2Weight d; memset(&d, 0, sizeof d); d.w0[0]=d.w0[1]=d.w0[2]=1; d.w1[0]=d.w1[1]=d.w1[2]=1; weights.push_back(d);
      WKey k; memcpy(k.v, &d, sizeof d); wmap[k] = 0; }

    int has_inter = 0, n_na = 0, first_na = nmb;
    const int save_addr = pic->CurrMbAddr, save_mbx = pic->mb_x, save_mby = pic->mb_y;

    for (int a = 0; a < nmb; a++) {
        CH264MacroBlock &mb = pic->m_mbs[a];
        H264B2MbInfo &I = info[a];
        coff[a] = (uint32_t)coefs.size();
        if (mb.m_name_of_mb_type == MB_TYPE_NA) { I.mb_class = H264B2_MB_NA; n_na++; if (a < first_na) first_na = a; continue; }
        int cls;
        if (mb.m_mb_pred_mode == Intra_4x4) cls = H264B2_MB_I4x4;
        else if (mb.m_mb_pred_mode == Intra_8x8) cls = H264B2_MB_I8x8;
        else if (mb.m_mb_pred_mode == Intra_16x16) cls = H264B2_MB_I16x16;
        else if (mb.m_name_of_mb_type == I_PCM) cls = H264B2_MB_IPCM;
        else cls = H264B2_MB_INTER;
        I.mb_class = (uint8_t)cls;
        I.flags = (mb.mb_field_decoding_flag ? H264B2_MBF_FIELD : 0) | (mb.transform_size_8x8_flag ? H264B2_MBF_T8x8 : 0)
                | ((mb.m_slice_type == H264_SLIECE_TYPE_SP || mb.m_slice_type == H264_SLIECE_TYPE_SI) ? H264B2_MBF_SPSI : 0)
                | ((IS_INTER_Prediction_Mode(mb.m_mb_pred_mode) && mb.constrained_intra_pred_flag == 1) ? H264B2_MBF_CIP_UNAVAIL : 0);
        I.pred16_chroma = (uint8_t)((mb.Intra16x16PredMode & 3) | ((mb.intra_chroma_pred_mode & 3) << 2));
        I.qpy = (int8_t)mb.QPY;
        if (mb.QP1Y != mb.QPY) { fprintf(stderr, "FATAL QP1Y != QPY (bit depth > 8?)\n"); exit(3); }
        I.slice_number = (uint16_t)mb.slice_number;
        uint16_t nnz = 0;
        for (int b = 0; b < 16; b++) {
            bool nz = mb.transform_size_8x8_flag ? (mb.mb_luma_8x8_non_zero_count_coeff[b >> 2] > 0)
                                                 : (mb.mb_luma_4x4_non_zero_count_coeff[b] > 0);
            if (nz) nnz |= (uint16_t)(1u << b);
        }
        I.nnz_mask = nnz;
        I.filter_offset_a = (int8_t)mb.FilterOffsetA; I.filter_offset_b = (int8_t)mb.FilterOffsetB;
        I.deblock_idc = (uint8_t)mb.disable_deblocking_filter_idc;
        if (mb.MbaffFrameFlag != mbaff) { fprintf(stderr, "FATAL per-MB MbaffFrameFlag differs from picture\n"); exit(3); }
        if (mb.TransformBypassModeFlag) { fprintf(stderr, "FATAL TransformBypassModeFlag unsupported\n"); exit(3); }

        // ---- coefficients (list order, before inverse scan) ----
        uint32_t cm = 0;
        if (cls == H264B2_MB_IPCM) {
            cm |= H264B2_CM_PCM;
            for (int i = 0; i < 256; i++) coefs.push_back((int16_t)mb.pcm_sample_luma[i]);
            for (int i = 0; i < 128; i++) coefs.push_back((int16_t)mb.pcm_sample_chroma[i]);
        } else {
            if (cls == H264B2_MB_I16x16) {
                for (int b = 0; b < 16; b++) if (anynz(mb.Intra16x16ACLevel[b], 15)) { cm |= H264B2_CM_LUMA(b); put16(coefs, mb.Intra16x16ACLevel[b], 16, 1); }
                if (anynz(mb.Intra16x16DCLevel, 16)) { cm |= H264B2_CM_LUMA_DC; put16(coefs, mb.Intra16x16DCLevel, 16, 0); }
            } else if (mb.transform_size_8x8_flag) {
                for (int b = 0; b < 4; b++) if (anynz(mb.LumaLevel8x8[b], 64)) { cm |= H264B2_CM_LUMA(b); put16(coefs, mb.LumaLevel8x8[b], 64, 0); }
            } else {
                for (int b = 0; b < 16; b++) if (anynz(mb.LumaLevel4x4[b], 16)) { cm |= H264B2_CM_LUMA(b); put16(coefs, mb.LumaLevel4x4[b], 16, 0); }
            }
            if (anynz(mb.ChromaDCLevel[0], 4) || anynz(mb.ChromaDCLevel[1], 4)) {
                cm |= H264B2_CM_CHROMA_DC; put16(coefs, mb.ChromaDCLevel[0], 4, 0); put16(coefs, mb.ChromaDCLevel[1], 4, 0);
            }
            for (int b = 0; b < 4; b++) if (anynz(mb.ChromaACLevel[0][b], 15)) { cm |= H264B2_CM_CB(b); put16(coefs, mb.ChromaACLevel[0][b], 16, 1); }
            for (int b = 0; b < 4; b++) if (anynz(mb.ChromaACLevel[1][b], 15)) { cm |= H264B2_CM_CR(b); put16(coefs, mb.ChromaACLevel[1][b], 16, 1); }
        }
        I.coef_mask = cm;

        if (cls == H264B2_MB_I4x4) { uint64_t m = 0; for (int b = 0; b < 16; b++) m |= (uint64_t)(mb.Intra4x4PredMode[b] & 15) << (4 * b); modes[a] = m; }
        if (cls == H264B2_MB_I8x8) { uint64_t m = 0; for (int b = 0; b < 4; b++) m |= (uint64_t)(mb.Intra8x8PredMode[b] & 15) << (4 * b); modes[a] = m; }

        if (cls != H264B2_MB_INTER) continue;
        has_inter = 1;
        H264B2MbMotion &M = mot[a];
        memset(M.ref_surf, -1, sizeof M.ref_surf); memset(M.ref_ident, -1, sizeof M.ref_ident);
        // reference state the member functions read
        pic->CurrMbAddr = a;
        pic->mb_x = (a % (W * (1 + mbaff))) / (1 + mbaff);
        pic->mb_y = (a / (W * (1 + mbaff)) * (1 + mbaff)) + ((a % (W * (1 + mbaff))) % (1 + mbaff));
        // the clamp in Luma_sample_interpolation_process reads m_h264_slice_data.mb_field_decoding_flag (IP:2351);
        // it is not recoverable per MB afterwards; the replay assumes it equals the MB's own flag.

        const bool direct16 = (mb.m_name_of_mb_type == B_Skip || mb.m_name_of_mb_type == B_Direct_16x16);
        const bool is8x8 = (mb.m_name_of_mb_type == P_8x8 || mb.m_name_of_mb_type == P_8x8ref0 || mb.m_name_of_mb_type == B_8x8);
        int NumMbPart = direct16 ? 4 : mb.m_NumMbPart;
        for (int mp = 0; mp < NumMbPart; mp++) {
            int nsub, pw, ph;
            if (!is8x8 && !direct16) { nsub = 1; pw = mb.MbPartWidth; ph = mb.MbPartHeight; }
            else if (mb.m_name_of_mb_type == P_8x8 || mb.m_name_of_mb_type == P_8x8ref0 ||
                     (mb.m_name_of_mb_type == B_8x8 && mb.m_name_of_sub_mb_type[mp] != B_Direct_8x8)) {
                nsub = mb.NumSubMbPart[mp]; pw = mb.SubMbPartWidth[mp]; ph = mb.SubMbPartHeight[mp];
            } else { nsub = 4; pw = 4; ph = 4; }
            int xP = (mp % (16 / mb.MbPartWidth)) * mb.MbPartWidth;
            int yP = (mp / (16 / mb.MbPartWidth)) * mb.MbPartHeight;
            int refIdx[2] = { mb.m_RefIdxL0[mp], mb.m_RefIdxL1[mp] };
            int pf[2] = { mb.m_PredFlagL0[mp], mb.m_PredFlagL1[mp] };
            // ---- reference surfaces, identities ----
            int8_t rs[2] = {-1, -1}, ri[2] = {-1, -1};
            for (int l = 0; l < 2; l++) {
                CH264Picture **list = l ? pic->m_RefPicList1 : pic->m_RefPicList0;
                int len = l ? pic->m_RefPicList1Length : pic->m_RefPicList0Length;
                if (pf[l]) {
                    CH264PictureBase *rp = nullptr;
                    int r = pic->Reference_picture_selection_process(refIdx[l], list, len, rp);
                    if (r != 0 || !rp || !rp->m_parent) { fprintf(stderr, "FATAL ref selection failed pic %d mb %d\n", g_decode_idx, a); exit(3); }
                    int s = slot_of(pic, rp->m_parent);
                    if (s < 0) { fprintf(stderr, "FATAL ref picture not in dpb\n"); exit(3); }
                    int view = (rp == &rp->m_parent->m_picture_frame) ? 0 : (rp == &rp->m_parent->m_picture_top_filed) ? 1 : 2;
                    rs[l] = (int8_t)((s << 2) | view);
                }
                if (refIdx[l] >= 0) {                 // DB:1175-1178: raw refIdx, pointer identity
                    if (refIdx[l] >= 16) { fprintf(stderr, "FATAL bS ref identity with refIdx >= 16\n"); exit(3); }
                    int s = slot_of(pic, list[refIdx[l]]);
                    if (s == -2) { fprintf(stderr, "FATAL stale ref pointer outside dpb\n"); exit(3); }
                    ri[l] = (int8_t)s;
                }
            }
            // ---- prediction weights (IP:538-546, 2545-2610), with the header of the slice this macroblock belongs to ----
            std::vector<WHdr> &hdrs = g_slice_hdrs[pic->m_parent];
            const bool swap_hdr = mb.slice_number >= 0 && (size_t)mb.slice_number < hdrs.size() && hdrs.size() > 1;
            if (swap_hdr && g_cur_hdr_slice != mb.slice_number) { if (g_cur_hdr_slice == -1) g_saved_hdr.load(pic->m_h264_slice_header); hdrs[mb.slice_number].store(pic->m_h264_slice_header); g_cur_hdr_slice = mb.slice_number; }
            int st = sh.slice_type % 5;
            int logWD[3] = {0,0,0}, w0[3] = {1,1,1}, w1[3] = {1,1,1}, o0[3] = {0,0,0}, o1[3] = {0,0,0};
            bool derive = (sh.m_pps.weighted_pred_flag == 1 && (st == 0 || st == 3)) || (sh.m_pps.weighted_bipred_idc > 0 && st == 1);
            if (derive) {
                int r = pic->Derivation_process_for_prediction_weights(refIdx[0], refIdx[1], pf[0], pf[1],
                        logWD[0], w0[0], w1[0], o0[0], o1[0], logWD[1], w0[1], w1[1], o0[1], o1[1], logWD[2], w0[2], w1[2], o0[2], o1[2]);
                if (r != 0) { fprintf(stderr, "FATAL weight derivation failed\n"); exit(3); }
            }
            int mode = 0;
            if (pf[0] == 1 && (st == 0 || st == 3)) mode = sh.m_pps.weighted_pred_flag ? 1 : 0;
            else if ((pf[0] || pf[1]) && st == 1) {
                if (sh.m_pps.weighted_bipred_idc == 1) mode = 1;
                else if (sh.m_pps.weighted_bipred_idc == 2) mode = (pf[0] && pf[1]) ? 1 : 0;
            }
            H264B2Weight we; memset(&we, 0, sizeof we);
            if (mode == 0) { we = weights[0]; }
            else {
                we.mode = 1;
                for (int c = 0; c < 3; c++) { we.logwd[c] = (int16_t)logWD[c]; we.w0[c] = (int16_t)w0[c]; we.w1[c] = (int16_t)w1[c]; we.o0[c] = (int16_t)o0[c]; we.o1[c] = (int16_t)o1[c]; }
                if (!pf[0] && pf[1]) { we.w1[0] = (int16_t)w0[0]; we.o1[0] = (int16_t)o0[0]; }   // Q8: IP:2764,2768 use w0L/o0L
                if (pf[0] && !pf[1]) { for (int c = 0; c < 3; c++) { we.w1[c] = 0; we.o1[c] = 0; } }          // canonicalise unused half
                if (!pf[0] && pf[1]) { for (int c = 0; c < 3; c++) { we.w0[c] = 0; we.o0[c] = 0; } }
            }
            WKey k; memset(&k, 0, sizeof k); memcpy(k.v, &we, sizeof we);
            int widx;
            auto it = wmap.find(k);
            if (it == wmap.end()) { widx = (int)weights.size(); weights.push_back(we); wmap[k] = widx; } else widx = it->second;
            if (widx > 65535) { fprintf(stderr, "FATAL too many weights\n"); exit(3); }
            // quadrants covered by this mb partition
            for (int qy = yP / 8; qy < (yP + mb.MbPartHeight + 7) / 8; qy++)
                for (int qx = xP / 8; qx < (xP + mb.MbPartWidth + 7) / 8; qx++) {
                    int q = qy * 2 + qx;
                    for (int l = 0; l < 2; l++) { M.ref_surf[l][q] = rs[l]; M.ref_ident[l][q] = ri[l]; }
                    M.wt_idx[q] = (uint16_t)widx;
                }
            for (int sp = 0; sp < nsub; sp++) {
                int xS, yS;
                if (is8x8) { xS = (sp % (8 / pw)) * pw; yS = (sp / (8 / pw)) * ph; }
                else { xS = (sp % 2) * 4; yS = (sp / 2) * 4; }
                for (int y = yP + yS; y < yP + yS + ph; y += 4)
                    for (int x = xP + xS; x < xP + xS + pw; x += 4) {
                        int r = (y / 4) * 4 + (x / 4);
                        M.mv[0][r][0] = (int16_t)mb.m_MvL0[mp][sp][0]; M.mv[0][r][1] = (int16_t)mb.m_MvL0[mp][sp][1];
                        M.mv[1][r][0] = (int16_t)mb.m_MvL1[mp][sp][0]; M.mv[1][r][1] = (int16_t)mb.m_MvL1[mp][sp][1];
                        if (mb.m_MvL0[mp][sp][0] != M.mv[0][r][0] || mb.m_MvL0[mp][sp][1] != M.mv[0][r][1] ||
                            mb.m_MvL1[mp][sp][0] != M.mv[1][r][0] || mb.m_MvL1[mp][sp][1] != M.mv[1][r][1]) { fprintf(stderr, "FATAL mv out of int16\n"); exit(3); }
                    }
            }
        }
    }
    pic->CurrMbAddr = save_addr; pic->mb_x = save_mbx; pic->mb_y = save_mby;
    if (g_cur_hdr_slice != -1) { g_saved_hdr.store(pic->m_h264_slice_header); g_cur_hdr_slice = -1; }
    g_slice_hdrs.erase(pic->m_parent);

    // ---- scaling: LevelScale in list order (PB:4852-4989 restated per scan position) ----
    bool flat = true;
    for (int l = 0; l < 6 && flat; l++) { for (int k = 0; k < 16; k++) if (sh.ScalingList4x4[l][k] != 16) flat = false; for (int k = 0; k < 64; k++) if (sh.ScalingList8x8[l][k] != 16) flat = false; }
    std::vector<int16_t> ls4, ls8;
    if (!flat) {
        for (int inter = 0; inter < 2; inter++) for (int fld = 0; fld < 2; fld++) for (int m = 0; m < 6; m++) {
            int32_t vals[16], c[4][4];
            for (int k = 0; k < 16; k++) vals[k] = k;
            pic->Inverse_scanning_process_for_4x4_transform_coefficients_and_scaling_lists(vals, c, fld);
            int pos[16][2]; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { pos[c[i][j]][0] = i; pos[c[i][j]][1] = j; }
            for (int k = 0; k < 16; k++) { int i = pos[k][0], j = pos[k][1];
                int n = (i % 2 == 0 && j % 2 == 0) ? norm4[m][0] : (i % 2 == 1 && j % 2 == 1) ? norm4[m][1] : norm4[m][2];
                ls4.push_back((int16_t)(sh.ScalingList4x4[inter ? 3 : 0][k] * n)); }
        }
        for (int inter = 0; inter < 2; inter++) for (int fld = 0; fld < 2; fld++) for (int m = 0; m < 6; m++) {
            int32_t vals[64], c[8][8];
            for (int k = 0; k < 64; k++) vals[k] = k;
            pic->Inverse_scanning_process_for_8x8_transform_coefficients_and_scaling_lists(vals, c, fld);
            int pos[64][2]; for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { pos[c[i][j]][0] = i; pos[c[i][j]][1] = j; }
            for (int k = 0; k < 64; k++) { int i = pos[k][0], j = pos[k][1]; int n;
                if (i % 4 == 0 && j % 4 == 0) n = norm8[m][0]; else if (i % 2 == 1 && j % 2 == 1) n = norm8[m][1];
                else if (i % 4 == 2 && j % 4 == 2) n = norm8[m][2]; else if ((i % 4 == 0 && j % 2 == 1) || (i % 2 == 1 && j % 4 == 0)) n = norm8[m][3];
                else if ((i % 4 == 0 && j % 4 == 2) || (i % 4 == 2 && j % 4 == 0)) n = norm8[m][4]; else n = norm8[m][5];
                ls8.push_back((int16_t)(sh.ScalingList8x8[inter ? 1 : 0][k] * n)); }
        }
    }
    while (coefs.size() % 4) coefs.push_back(0);

    PicHdr h; memset(&h, 0, sizeof h);
    h.decode_idx = g_decode_idx; h.dst_surface = slot_of(pic, pic->m_parent);
    h.clear_surface = n_na > 0; h.has_inter = has_inter; h.deblock_enable = deblock_enable; h.deblock_stop_mb = first_na;
    h.mbaff = mbaff; h.cqp0 = sh.m_pps.chroma_qp_index_offset; h.cqp1 = sh.m_pps.second_chroma_qp_index_offset;
    h.n_weights = (int)weights.size(); h.custom_scaling = !flat; h.slice_type = sh.slice_type; h.poc = pic->PicOrderCnt; h.n_na = n_na;
    h.n_coefs = (uint32_t)coefs.size(); h.nal_ref_idc = sh.m_nal_unit.nal_ref_idc; h.sum_pre = sum_pre; h.sum_post = sum_post;
    if (h.dst_surface < 0) { fprintf(stderr, "FATAL current picture not in dpb\n"); exit(3); }
    fwrite(&h, sizeof h, 1, g_replay);
    fwrite(info.data(), sizeof(H264B2MbInfo), nmb, g_replay);
    fwrite(modes.data(), 8, nmb, g_replay);
    fwrite(coff.data(), 4, nmb, g_replay);
    if (has_inter) fwrite(mot.data(), sizeof(H264B2MbMotion), nmb, g_replay);
    fwrite(weights.data(), sizeof(H264B2Weight), weights.size(), g_replay);
    fwrite(coefs.data(), 2, coefs.size(), g_replay);
    if (!flat) { fwrite(ls4.data(), 2, ls4.size(), g_replay); fwrite(ls8.data(), 2, ls8.size(), g_replay); }
}

static void write_pix(CH264PictureBase *f, const char *tag) {
    if (!g_pixdump) return;
    (void)tag;
    fwrite(f->m_pic_buff_luma, 1, (size_t)f->PicWidthInSamplesL * f->PicHeightInSamplesL, g_pixdump);
    fwrite(f->m_pic_buff_cb, 1, (size_t)f->PicWidthInSamplesC * f->PicHeightInSamplesC, g_pixdump);
    fwrite(f->m_pic_buff_cr, 1, (size_t)f->PicWidthInSamplesC * f->PicHeightInSamplesC, g_pixdump);
}

// ---------------------------------------------------------------- link-time interposers
extern "C" int __real__ZN16CH264PictureBase25Deblocking_filter_processEv(CH264PictureBase *self);
extern "C" int __wrap__ZN16CH264PictureBase25Deblocking_filter_processEv(CH264PictureBase *self) {
    if (!g_replay) { g_decode_idx++; return __real__ZN16CH264PictureBase25Deblocking_filter_processEv(self); }
    uint64_t pre = checksum_pic(*self);
    // snapshot what dump_picture needs BEFORE deblocking only for pixels; the MB state is not modified by deblocking
    bool dumppix = (g_dump_pix_idx == g_decode_idx);
    if (dumppix) write_pix(self, "pre");
    int r = __real__ZN16CH264PictureBase25Deblocking_filter_processEv(self);
    uint64_t post = checksum_pic(*self);
    if (dumppix) write_pix(self, "post");
    dump_picture(self, 1, pre, post);
    g_pic2idx[self->m_parent] = g_decode_idx;
    g_last_dumped = self->m_parent; g_last_dumped_numcnt = self->m_PicNumCnt;
    g_decode_idx++;
    return r;
}
extern "C" int __real__ZN12CH264Picture16decode_one_sliceER16CH264SliceHeaderR10CBitstreamRA16_PS_(CH264Picture *, CH264SliceHeader *, CBitstream *, CH264Picture **);
extern "C" int __wrap__ZN12CH264Picture16decode_one_sliceER16CH264SliceHeaderR10CBitstreamRA16_PS_(CH264Picture *self, CH264SliceHeader *sh, CBitstream *bs, CH264Picture **dpb) {
    g_cur_pic = self;
    if (g_replay) { WHdr w; w.load(*sh); g_slice_hdrs[self].push_back(w); }
    return __real__ZN12CH264Picture16decode_one_sliceER16CH264SliceHeaderR10CBitstreamRA16_PS_(self, sh, bs, dpb);
}

static void dump_tail_picture() {   // the last picture in decoding order is never deblocked (VD:354-361, Q1)
    if (!g_replay || !g_cur_pic || !g_cur_pic->m_current_picture_ptr) return;
    CH264PictureBase *f = g_cur_pic->m_current_picture_ptr;
    if (g_cur_pic == g_last_dumped && f->m_PicNumCnt == g_last_dumped_numcnt) return;
    uint64_t s = checksum_pic(*f);
    if (g_dump_pix_idx == g_decode_idx) { write_pix(f, "pre"); write_pix(f, "post"); }
    dump_picture(f, 0, s, s);
    g_pic2idx[g_cur_pic] = g_decode_idx;
    g_last_dumped = g_cur_pic; g_last_dumped_numcnt = f->m_PicNumCnt;
    g_decode_idx++;
}

static int cb(CH264Picture *pic, void *, int) {
    if (!pic) { dump_tail_picture(); return -1; }   // FILE_END marker: pictures_gop is still alive here (VD:372-377)
    CH264PictureBase &f = pic->m_picture_frame;
    // a picture that reaches the callback before its (never executed) deblocking is the tail picture
    if (g_replay && pic == g_cur_pic && !(pic == g_last_dumped && f.m_PicNumCnt == g_last_dumped_numcnt)) dump_tail_picture();
    uint64_t h = checksum_pic(f);
    g_stream_hash = g_stream_hash * 0x100000001B3ULL + h;
    if (!g_quiet) fprintf(stderr, "frame %d %dx%d poc=%d type=%d sum=%016llx\n", g_nframes, f.PicWidthInSamplesL, f.PicHeightInSamplesL,
                          f.PicOrderCnt, f.m_h264_slice_header.slice_type, (unsigned long long)h);
    if (g_yuv) {
        fwrite(f.m_pic_buff_luma, 1, (size_t)f.PicWidthInSamplesL * f.PicHeightInSamplesL, g_yuv);
        fwrite(f.m_pic_buff_cb, 1, (size_t)f.PicWidthInSamplesC * f.PicHeightInSamplesC, g_yuv);
        fwrite(f.m_pic_buff_cr, 1, (size_t)f.PicWidthInSamplesC * f.PicHeightInSamplesC, g_yuv);
    }
    if (g_bgr) {    // the reference's own YUV420P -> BGR24 conversion (H264PictureBase.cpp:440), checksummed per output frame
        const int W = f.PicWidthInSamplesL, H = f.PicHeightInSamplesL;
        std::vector<uint8_t> bgr((size_t)W * 3 * H);
        f.convertYuv420pToBgr24(W, H, f.m_pic_buff_luma, bgr.data(), W * 3);
        fprintf(g_bgr, "%d %d %016llx %016llx\n", g_nframes, f.PicOrderCnt, (unsigned long long)h, (unsigned long long)checksum_bytes(bgr.data(), bgr.size(), 0, 0));
    }
    if (g_replay) {
        auto it = g_pic2idx.find(pic);
        OutRec o; o.decode_idx = (it == g_pic2idx.end()) ? -1 : it->second; o.pad = 0; o.sum = h; g_out.push_back(o);
    }
    return (++g_nframes >= g_maxframes) ? -1 : 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s in.h264 [--yuv F] [--replay F] [--max-frames N] [--quiet] [--bgr-sums F] [--dump-pix IDX F]\n", argv[0]); return 2; }
    const char *replay_path = nullptr;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "--yuv") && i + 1 < argc) g_yuv = fopen(argv[++i], "wb");
        else if (!strcmp(argv[i], "--replay") && i + 1 < argc) { replay_path = argv[++i]; g_replay = fopen(replay_path, "wb"); }
        else if (!strcmp(argv[i], "--max-frames") && i + 1 < argc) g_maxframes = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) g_quiet = 1;
        else if (!strcmp(argv[i], "--bgr-sums") && i + 1 < argc) g_bgr = fopen(argv[++i], "w");
        else if (!strcmp(argv[i], "--dump-pix") && i + 2 < argc) { g_dump_pix_idx = atol(argv[++i]); g_pixdump = fopen(argv[++i], "wb"); }
    }
    struct FileHdr { char magic[8]; uint32_t version, width_mbs, height_mbs, n_pics, n_out, hdr_bytes, pichdr_bytes, reserved; } fh;
    if (g_replay) { memset(&fh, 0, sizeof fh); fwrite(&fh, sizeof fh, 1, g_replay); }
    CH264VideoDecoder vd;
    vd.set_output_frame_callback_functuin(cb, nullptr);
    auto t0 = std::chrono::steady_clock::now();
    int r = vd.open(argv[1]);
    auto t1 = std::chrono::steady_clock::now();
    double s = std::chrono::duration<double>(t1 - t0).count();
    if (g_replay) {
        fwrite(g_out.data(), sizeof(OutRec), g_out.size(), g_replay);
        memcpy(fh.magic, "H264B2RP", 8); fh.version = 1; fh.width_mbs = g_wmb; fh.height_mbs = g_hmb; fh.n_pics = g_decode_idx; fh.n_out = (uint32_t)g_out.size();
        fh.hdr_bytes = sizeof fh; fh.pichdr_bytes = sizeof(PicHdr);
        fseek(g_replay, 0, SEEK_SET); fwrite(&fh, sizeof fh, 1, g_replay); fclose(g_replay);
    }
    if (g_yuv) fclose(g_yuv);
    if (g_pixdump) fclose(g_pixdump);
    if (g_bgr) fclose(g_bgr);
    fprintf(stderr, "RESULT file=%s ret=%d frames=%d pics=%d secs=%.3f fps=%.3f stream_hash=%016llx\n", argv[1], r, g_nframes, g_decode_idx, s, g_nframes / s,
            (unsigned long long)g_stream_hash);
    return 0;
}
