// h264_slice.cpp — slice_data(), macroblock_layer(), residual (CAVLC + CABAC) and every per-macroblock derivation the
// reference performs inside its reconstruction calls, restated for the SoA output:
//   slice loop            H264SliceData.cpp:64-534          macroblock syntax     H264MacroBlock.cpp:912-1856
//   CAVLC residual        H264ResidualBlockCavlc.cpp:35-578  CABAC syntax          H264Cabac.cpp:1088-5161
//   neighbours (6.4.11/12) H264PictureBase.cpp:2503-3395     intra pred modes      H264PictureBase.cpp:773-1060
//   motion vectors        H264InterPrediction.cpp:412-2047   weights / ref select  H264InterPrediction.cpp:2117-3047
// Deviations of the reference from H.264 that change results are kept and marked "REF:".
#include "h264_decoder.h"
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <emmintrin.h>      // SSE2 is part of the x86-64 baseline

// -DFE_PROF: rdtsc section counters (single-threaded profiling builds only; printed at exit).  The sections nest: the numbers are
// for finding the expensive stage, each rdtsc pair costs ~50 cycles itself.
#ifdef FE_PROF
#include <x86intrin.h>
unsigned long long g_prof[16];
struct ProfDump { ~ProfDump() { const char *n[] = {"mb_layer(all)", "residual", "inter_pred", "intra_modes", "emit_mb", "skip_mb", "slice_total", "cbf_ctx", "cbf+sigmap", "mb_type", "pred_syntax", "cbp+qpd", "skipflag+term"}; for (int i = 0; i < 13; i++) fprintf(stderr, "%-14s %8.1f Mcycles\n", n[i], g_prof[i] / 1e6); } } g_profdump;
#define PROF_T(v) const unsigned long long v = __rdtsc()
#define PROF_ADD(i, v) g_prof[i] += __rdtsc() - v
#else
#define PROF_T(v)
#define PROF_ADD(i, v)
#endif
namespace h264b2 {

namespace {

const uint8_t kBlkX[16] = {0, 4, 0, 4, 8, 12, 8, 12, 0, 4, 0, 4, 8, 12, 8, 12};     // luma4x4BlkIdx -> x, y (6.4.3)
const uint8_t kBlkY[16] = {0, 0, 4, 4, 0, 0, 4, 4, 8, 8, 12, 12, 8, 8, 12, 12};
inline int blk_of_xy(int x, int y) { return 8 * (y / 8) + 4 * (x / 8) + 2 * ((y % 8) / 4) + ((x % 8) / 4); }

// Table 9-43 ctxIdxInc for 8x8 blocks: significant_coeff_flag (frame, field coded), last_significant_coeff_flag
const uint8_t kSig8Frame[63] = {0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7, 7, 6, 11, 12, 13, 11, 6, 7, 8, 9, 14, 10, 9, 8, 6, 11,
                                12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12};
const uint8_t kSig8Field[63] = {0, 1, 1, 2, 2, 3, 3, 4, 5, 6, 7, 7, 7, 8, 4, 5, 6, 9, 10, 10, 8, 11, 12, 11, 9, 9, 10, 10, 8, 11, 12, 11, 9, 9, 10, 10, 8, 11, 12, 11, 9, 9, 10, 10, 8, 13, 13, 9,
                                9, 10, 10, 8, 13, 13, 9, 9, 10, 10, 14, 14, 14, 14, 14};
const uint8_t kLast8[63] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8};

enum { CAT_I16DC = 0, CAT_I16AC = 1, CAT_LUMA4 = 2, CAT_CDC = 3, CAT_CAC = 4, CAT_LUMA8 = 5 };

struct Dec {
    Front &F; Slot &S; const SliceHeader &sh; BitReader &br; Cabac &cb;
    MbT *mbs; H264B2MbMotion *mot; const int W, nmb; const bool cabac; const int mbaff;
    int cur = 0;                      // CurrMbAddr
    // residual of the macroblock being parsed (list order, as the reference's level arrays)
    int32_t i16dc[16], i16ac[16][16], l4[16][16], l8[4][64], cdc[2][4], cac[2][4][16];
    int16_t pcm[384];
    bool residual_ok = false;
    Dec(Front &f) : F(f), S(f.slots[f.cur]), sh(f.sh), br(f.br), cb(f.cabac), mbs(f.mbs.data()), mot(S.motion.data()), W(f.wmb), nmb(f.nmb),
                    cabac(f.sh.pps.entropy_coding_mode_flag != 0), mbaff(f.sh.MbaffFrameFlag) {}

    // ------------------------------------------------------------ neighbouring locations (6.4.12; PB:2878-3395)
    inline bool same_slice(int a, int c) const { return mbs[a].slice == mbs[c].slice; }
    // non-MBAFF: the four neighbouring macroblock addresses of the current macroblock are derived once per macroblock
    mutable int nc_mb = -1, nc_slice = -1, nA_ = -1, nB_ = -1, nC_ = -1, nD_ = -1;
    inline void nbr_cache(int c) const {
        nc_mb = c; nc_slice = mbs[c].slice;
        const int col = c % W;
        int n;
        n = c - 1; nA_ = (n < 0 || col == 0 || !same_slice(n, c)) ? -1 : n;
        n = c - W; nB_ = (n < 0 || !same_slice(n, c)) ? -1 : n;
        n = c - W + 1; nC_ = (n < 0 || col == W - 1 || !same_slice(n, c)) ? -1 : n;
        n = c - W - 1; nD_ = (n < 0 || col == 0 || !same_slice(n, c)) ? -1 : n;
    }
    // progressive pictures: four cached addresses and two masks — kept inline, it is called ~30 times per macroblock
    __attribute__((always_inline)) inline Nb nbr(int c, int xN, int yN, bool chroma) const {
        if (mbaff) return nbr_mbaff(c, xN, yN, chroma);
        const int maxW = chroma ? 8 : 16, maxH = chroma ? 8 : 16;
        Nb r; r.mb = -1; r.xW = 0; r.yW = 0;
        if (yN > maxH - 1) return r;
        if (c != nc_mb || mbs[c].slice != nc_slice) nbr_cache(c);
        int n;
        if (xN < 0) n = yN < 0 ? nD_ : nA_;
        else if (xN <= maxW - 1) n = yN < 0 ? nB_ : c;
        else n = yN < 0 ? nC_ : -1;
        r.mb = n; r.xW = (xN + maxW) & (maxW - 1); r.yW = (yN + maxH) & (maxH - 1);
        return r;
    }
    __attribute__((noinline)) Nb nbr_mbaff(int c, int xN, int yN, bool chroma) const {
        const int maxW = chroma ? 8 : 16, maxH = chroma ? 8 : 16;
        Nb r; r.mb = -1; r.xW = 0; r.yW = 0;
        if (yN > maxH - 1) return r;
        // MBAFF (Table 6-4)
        int A = 2 * (c / 2 - 1), B = 2 * (c / 2 - W), C = 2 * (c / 2 - W + 1), D = 2 * (c / 2 - W - 1);
        if (A < 0 || A > c || !same_slice(A, c) || (c / 2) % W == 0) A = -2;
        if (B < 0 || B > c || !same_slice(B, c)) B = -2;
        if (C < 0 || C > c || !same_slice(C, c) || (c / 2 + 1) % W == 0) C = -2;
        if (D < 0 || D > c || !same_slice(D, c) || (c / 2) % W == 0) D = -2;
        const bool curFrame = mbs[c].field == 0, top = (c % 2) == 0;
        int n = -1, yM = 0;
        auto frameX = [&](int x) { return mbs[x].field == 0; };
        if (xN < 0 && yN < 0) {
            if (curFrame) {
                if (top) { n = D + 1; yM = yN; if (D < 0) n = -1; }
                else if (A >= 0) { if (frameX(A)) { n = A; yM = yN; } else { n = A + 1; yM = (yN + maxH) >> 1; } }      // REF/spec: bottom frame MB, left field pair -> bottom field MB
            } else {
                if (top) { if (D >= 0) { if (frameX(D)) { n = D + 1; yM = 2 * yN; } else { n = D; yM = yN; } } }
                else { n = D + 1; yM = yN; if (D < 0) n = -1; }
            }
        } else if (xN < 0) {
            if (A >= 0) {
                if (curFrame) {
                    if (top) { if (frameX(A)) { n = A; yM = yN; } else { n = (yN % 2 == 0) ? A : A + 1; yM = yN >> 1; } }
                    else { if (frameX(A)) { n = A + 1; yM = yN; } else { n = (yN % 2 == 0) ? A : A + 1; yM = (yN + maxH) >> 1; } }
                } else {
                    if (top) { if (frameX(A)) { if (yN < maxH / 2) { n = A; yM = yN << 1; } else { n = A + 1; yM = (yN << 1) - maxH; } } else { n = A; yM = yN; } }
                    else { if (frameX(A)) { if (yN < maxH / 2) { n = A; yM = (yN << 1) + 1; } else { n = A + 1; yM = (yN << 1) + 1 - maxH; } } else { n = A + 1; yM = yN; } }
                }
            }
        } else if (xN <= maxW - 1 && yN < 0) {
            if (curFrame) {
                if (top) { n = B + 1; yM = yN; if (B < 0) n = -1; }
                else { n = c - 1; yM = yN; }
            } else {
                if (top) { if (B >= 0) { if (frameX(B)) { n = B + 1; yM = 2 * yN; } else { n = B; yM = yN; } } }
                else { n = B + 1; yM = yN; if (B < 0) n = -1; }
            }
        } else if (xN <= maxW - 1) { n = c; yM = yN; }
        else if (yN < 0) {
            if (curFrame) {
                if (top) { n = C + 1; yM = yN; if (C < 0) n = -1; }
            } else {
                if (top) { if (C >= 0) { if (frameX(C)) { n = C + 1; yM = 2 * yN; } else { n = C; yM = yN; } } }
                else { n = C + 1; yM = yN; if (C < 0) n = -1; }
            }
        }
        if (n < 0) { r.mb = -1; return r; }
        r.mb = n; r.xW = (xN + maxW) % maxW; r.yW = (yM + maxH) % maxH;
        return r;
    }
    inline int nbrA(int c) const { return nbr(c, -1, 0, false).mb; }      // 6.4.11.1
    inline int nbrB(int c) const { return nbr(c, 0, -1, false).mb; }

    static inline bool is_skip(uint8_t t) { return t == T_PSKIP || t == T_BSKIP; }

    // ------------------------------------------------------------ CABAC syntax elements
    int cabac_mb_skip_flag(int addr) {
        const int A = nbrA(addr), B = nbrB(addr);
        const int inc = (A >= 0 && !mbs[A].skip_flag) + (B >= 0 && !mbs[B].skip_flag);
        return cb.decision(((sh.slice_type == SLICE_B) ? 24 : 11) + inc);
    }
    int cabac_mb_field_flag() {
        int A = 2 * (cur / 2 - 1), B = 2 * (cur / 2 - W);
        if (A < 0 || !same_slice(A, cur) || (cur / 2) % W == 0) A = -1;
        if (B < 0 || !same_slice(B, cur)) B = -1;
        const int inc = (A >= 0 && mbs[A].field) + (B >= 0 && mbs[B].field);
        return cb.decision(70 + inc);
    }
    int cabac_intra_mb_type(int base, bool islice) {       // returns I-slice mb_type 0..25
        if (islice) {
            const int A = nbrA(cur), B = nbrB(cur);
            const int inc = (A >= 0 && mbs[A].type != T_I_NxN) + (B >= 0 && mbs[B].type != T_I_NxN);
            if (!cb.decision(base + inc)) return 0;
            if (cb.terminate()) { cb.init_engine(&br); return 25; }      // REF: see macroblock_layer (I_PCM)
            int t = 1 + 12 * cb.decision(base + 3);
            if (cb.decision(base + 4)) t += 4 + 4 * cb.decision(base + 5);
            t += 2 * cb.decision(base + 6);
            t += cb.decision(base + 7);
            return t;
        }
        if (!cb.decision(base)) return 0;
        if (cb.terminate()) { cb.init_engine(&br); return 25; }
        int t = 1 + 12 * cb.decision(base + 1);
        if (cb.decision(base + 2)) t += 4 + 4 * cb.decision(base + 2);
        t += 2 * cb.decision(base + 3);
        t += cb.decision(base + 3);
        return t;
    }
    int cabac_mb_type() {
        const int st = sh.slice_type;
        if (st == SLICE_I) return cabac_intra_mb_type(3, true);
        if (st == SLICE_P || st == SLICE_SP) {
            if (!cb.decision(14)) {
                if (!cb.decision(15)) return 3 * cb.decision(16);
                return 2 - cb.decision(17);
            }
            return 5 + cabac_intra_mb_type(17, false);
        }
        if (st == SLICE_B) {
            const int A = nbrA(cur), B = nbrB(cur);
            const int inc = (A >= 0 && mbs[A].type != T_BSKIP && mbs[A].type != T_BDIRECT) + (B >= 0 && mbs[B].type != T_BSKIP && mbs[B].type != T_BDIRECT);
            if (!cb.decision(27 + inc)) return 0;
            if (!cb.decision(27 + 3)) return 1 + cb.decision(27 + 5);
            int bits = cb.decision(27 + 4) << 3;
            bits |= cb.decision(27 + 5) << 2; bits |= cb.decision(27 + 5) << 1; bits |= cb.decision(27 + 5);
            if (bits < 8) return bits + 3;
            if (bits == 13) return 23 + cabac_intra_mb_type(32, false);
            if (bits == 14) return 11;
            if (bits == 15) return 22;
            bits = (bits << 1) | cb.decision(27 + 5);
            return bits - 4;
        }
        return -1;     // SI slices are not supported
    }
    int cabac_sub_mb_type_p() { if (cb.decision(21)) return 0; if (!cb.decision(22)) return 1; return cb.decision(23) ? 2 : 3; }
    int cabac_sub_mb_type_b() {
        if (!cb.decision(36)) return 0;
        if (!cb.decision(37)) return 1 + cb.decision(39);
        int t = 3;
        if (cb.decision(38)) { if (cb.decision(39)) return 11 + cb.decision(39); t += 4; }
        t += 2 * cb.decision(39); t += cb.decision(39);
        return t;
    }
    int cabac_t8x8_flag() {
        const int A = nbrA(cur), B = nbrB(cur);
        return cb.decision(399 + (A >= 0 && mbs[A].t8x8) + (B >= 0 && mbs[B].t8x8));
    }
    int cabac_intra_chroma_pred_mode() {
        const int A = nbrA(cur), B = nbrB(cur);
        const int inc = ctx_chroma(A) + ctx_chroma(B);
        if (!cb.decision(64 + inc)) return 0;
        if (!cb.decision(64 + 3)) return 1;
        return cb.decision(64 + 3) ? 3 : 2;
    }
    int ctx_chroma(int n) const {       // 9.3.3.1.1.8: inter MB, I_PCM or intra_chroma_pred_mode == 0 -> 0
        if (n < 0) return 0;
        const MbT &m = mbs[n];
        if (!(m.type == T_I_NxN || m.type == T_I16)) return 0;
        return m.chroma_pred != 0;
    }
    int cabac_cbp() {
        int luma = 0;
        for (int b8 = 0; b8 < 4; b8++) {
            const int x = (b8 % 2) * 8, y = (b8 / 2) * 8;
            int cond[2];
            for (int k = 0; k < 2; k++) {
                const Nb n = k == 0 ? nbr(cur, x - 1, y, false) : nbr(cur, x, y - 1, false);
                int c;
                if (n.mb < 0) c = 0;
                else if (n.mb == cur) { const int b8n = (n.yW / 8) * 2 + n.xW / 8; c = ((luma >> b8n) & 1) != 0 ? 0 : 1; }
                else {
                    const MbT &m = mbs[n.mb];
                    const int b8n = (n.yW / 8) * 2 + n.xW / 8;
                    if (m.type == T_IPCM) c = 0;
                    else if (!is_skip(m.type) && ((m.cbp_luma >> b8n) & 1) != 0) c = 0;
                    else c = 1;
                }
                cond[k] = c;
            }
            luma |= cb.decision(73 + cond[0] + 2 * cond[1]) << b8;
        }
        const int A = nbrA(cur), B = nbrB(cur);
        auto cc = [&](int n, int bin) {
            if (n < 0) return 0;
            const MbT &m = mbs[n];
            if (m.type == T_IPCM) return 1;
            if (is_skip(m.type)) return 0;
            if (bin == 0) return m.cbp_chroma != 0 ? 1 : 0;
            return m.cbp_chroma == 2 ? 1 : 0;
        };
        int chroma = 0;
        if (cb.decision(77 + cc(A, 0) + 2 * cc(B, 0))) chroma = 1 + cb.decision(77 + 4 + cc(A, 1) + 2 * cc(B, 1));
        return luma | (chroma << 4);
    }
    int cabac_mb_qp_delta() {
        int prev = cur - 1;
        if (cur == sh.first_mb_in_slice * (1 + mbaff)) prev = -1;
        int inc = 1;
        if (prev < 0) inc = 0;
        else {
            const MbT &m = mbs[prev];
            if (is_skip(m.type) || m.type == T_IPCM || (m.type != T_I16 && m.cbp_luma == 0 && m.cbp_chroma == 0) || m.qp_delta == 0) inc = 0;
        }
        if (!cb.decision(60 + inc)) return 0;
        int bin = cb.decision(60 + 2), binIdx = 1;
        while (bin) { bin = cb.decision(60 + 3); binIdx++; if (binIdx > 102) return QP_DELTA_ERROR; }      // H264Cabac.cpp:4196: "too large" fails the macroblock
        return (binIdx & 1) ? (binIdx + 1) >> 1 : -((binIdx + 1) >> 1);
    }
    enum { QP_DELTA_ERROR = 0x7fffffff };
    // neighbouring partition of the current MB's partition at (x, y): availability, quadrant and 4x4 block in the neighbour
    // REF: both context derivations look the neighbour's partition mode up with MbPartPredMode2(), which knows no I_PCM
    // (H264MacroBlock.cpp:756): an available I_PCM neighbour FAILS the syntax element and with it the macroblock (H264Cabac.cpp:1466, 1736)
    enum { MVD_ERROR = 0x7ffffffe };
    int cabac_ref_idx(int list, int x, int y) {
        int cond[2];
        for (int k = 0; k < 2; k++) {
            const Nb n = k == 0 ? nbr(cur, x - 1, y, false) : nbr(cur, x, y - 1, false);
            int c = 0;
            if (n.mb >= 0) {
                const MbT &m = mbs[n.mb];
                if (m.type == T_IPCM) return -1;
                const int q = (n.yW / 8) * 2 + n.xW / 8;
                const int thr = (mbaff && mbs[cur].field == 0 && m.field == 1) ? 1 : 0;
                const bool zero = !(m.ref_syn[list][q] > thr);
                const bool eq = (m.part_pm[q] & (1 << list)) != 0 && m.part_pm[q] != PM_DIRECT;
                c = !(is_skip(m.type) || m.intra || m.type == T_IPCM || !eq || zero);
            }
            cond[k] = c;
        }
        if (!cb.decision(54 + cond[0] + 2 * cond[1])) return 0;
        int bin = cb.decision(54 + 4), binIdx = 1;
        while (bin) { bin = cb.decision(54 + 5); binIdx++; if (binIdx > 32) return -1; }      // H264Cabac.cpp:4133
        return binIdx;
    }
    int cabac_mvd(int list, int comp, int x, int y) {
        int sum = 0;
        for (int k = 0; k < 2; k++) {
            const Nb n = k == 0 ? nbr(cur, x - 1, y, false) : nbr(cur, x, y - 1, false);
            if (n.mb < 0) continue;
            const MbT &m = mbs[n.mb];
            if (m.type == T_IPCM) return MVD_ERROR;
            const int q = (n.yW / 8) * 2 + n.xW / 8, b = (n.yW / 4) * 4 + n.xW / 4;
            const bool eq = (m.part_pm[q] & (1 << list)) != 0 && m.part_pm[q] != PM_DIRECT;
            if (is_skip(m.type) || m.intra || !eq) continue;
            int a = abs(m.mvd[list][b][comp]);
            if (comp == 1 && mbaff) { if (mbs[cur].field == 0 && m.field == 1) a *= 2; else if (mbs[cur].field == 1 && m.field == 0) a /= 2; }
            sum += a;
        }
        const int base = comp ? 47 : 40;
        const int inc = sum < 3 ? 0 : sum > 32 ? 2 : 1;
        if (!cb.decision(base + inc)) return 0;
        int v = 1, ctx = base + 3;
        while (v < 9) { if (!cb.decision(ctx)) break; v++; if (v <= 4) ctx++; }
        if (v >= 9) {
            int k = 3;
            while (cb.bypass()) { v += 1 << k; k++; if (k >= 23) return v; }      // REF: H264Cabac.cpp:4048 returns "success" here: no suffix, no sign
            while (k--) v += cb.bypass() << k;
        }
        return cb.bypass() ? -v : v;
    }
    // coded_block_flag context (9.3.3.1.1.9; H264Cabac.cpp:1996-2526)
    int cbf_ctx(int cat, int blk, int comp /* -1 luma, 0 Cb, 1 Cr */) {
        int cond[2];
        const MbT &cm = mbs[cur];
        for (int k = 0; k < 2; k++) {
            int nmbk = -1, avail_tb = 0, flag = 0;
            if (cat == CAT_I16DC || cat == CAT_CDC) {
                nmbk = k == 0 ? nbr(cur, -1, 0, cat == CAT_CDC).mb : nbr(cur, 0, -1, cat == CAT_CDC).mb;
                if (nmbk >= 0) {
                    const MbT &m = mbs[nmbk];
                    if (cat == CAT_I16DC) { if (m.type == T_I16) { avail_tb = 1; flag = m.cbf_dc & 1; } }
                    else if (!is_skip(m.type) && m.type != T_IPCM && m.cbp_chroma != 0) { avail_tb = 1; flag = (m.cbf_dc >> (comp + 1)) & 1; }
                }
            } else if (cat == CAT_I16AC || cat == CAT_LUMA4) {
                const int x = kBlkX[blk], y = kBlkY[blk];
                const Nb n = k == 0 ? nbr(cur, x - 1, y, false) : nbr(cur, x, y - 1, false);
                nmbk = n.mb;
                if (nmbk >= 0) {
                    const MbT &m = mbs[nmbk];
                    const int b4 = blk_of_xy(n.xW, n.yW);
                    if (!is_skip(m.type) && ((m.cbp_luma >> (b4 >> 2)) & 1)) {
                        if (m.t8x8 == 0) { if (m.type != T_IPCM) { avail_tb = 1; flag = (m.cbf_ac[0] >> b4) & 1; } }
                        else { avail_tb = 1; flag = (m.cbf_ac[0] >> (b4 >> 2)) & 1; }
                    }
                }
            } else if (cat == CAT_CAC) {
                const int x = (blk % 2) * 4, y = (blk / 2) * 4;
                const Nb n = k == 0 ? nbr(cur, x - 1, y, true) : nbr(cur, x, y - 1, true);
                nmbk = n.mb;
                if (nmbk >= 0) {
                    const MbT &m = mbs[nmbk];
                    const int b4 = 2 * (n.yW / 4) + n.xW / 4;
                    if (!is_skip(m.type) && m.type != T_IPCM && m.cbp_chroma == 2) { avail_tb = 1; flag = (m.cbf_ac[comp + 1] >> b4) & 1; }
                }
            }
            int c;
            if ((nmbk < 0 && !cm.intra) || (nmbk >= 0 && !avail_tb && mbs[nmbk].type != T_IPCM)) c = 0;
            else if (nmbk < 0 || mbs[nmbk].type == T_IPCM) c = 1;
            else c = flag;
            cond[k] = c;
        }
        return cond[0] + 2 * cond[1];
    }
    int residual_block_cabac(int32_t *lvl, int startIdx, int endIdx, int maxNumCoeff, int cat, int blk, int comp) {
        static const int cbfOff[5] = {0, 4, 8, 12, 16}, sigOff[5] = {0, 15, 29, 44, 47}, absOff[5] = {0, 10, 20, 30, 39};
        MbT &m = mbs[cur];
        int coded = 1;
        PROF_T(tc);
        if (maxNumCoeff != 64) coded = cb.decision(85 + cbfOff[cat] + cbf_ctx(cat, blk, comp));
        PROF_ADD(7, tc);
        memset(lvl, 0, sizeof(int32_t) * maxNumCoeff);
        if (!coded) return 0;
        const int fld = m.field;
        int sigBase, lastBase, absBase;
        if (cat == CAT_LUMA8) { sigBase = fld ? 436 : 402; lastBase = fld ? 451 : 417; absBase = 426; }
        else { sigBase = (fld ? 277 : 105) + sigOff[cat]; lastBase = (fld ? 338 : 166) + sigOff[cat]; absBase = 227 + absOff[cat]; }
        static const uint8_t ident[64] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41,
                                          42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63};
        static const uint8_t cdcInc[4] = {0, 1, 2, 2};
        const uint8_t *sigInc = cat == CAT_LUMA8 ? (fld ? kSig8Field : kSig8Frame) : cat == CAT_CDC ? cdcInc : ident;
        const uint8_t *lastInc = cat == CAT_LUMA8 ? kLast8 : cat == CAT_CDC ? cdcInc : ident;
        uint8_t where[64];            // positions of the significant coefficients, in scan order
        int nsig = 0, numCoeff = endIdx + 1, i = startIdx;
        {
        CabacRegs r(cb);
        while (i < numCoeff - 1) {
            if (r.decision(sigBase + sigInc[i])) {
                where[nsig++] = (uint8_t)i;
                if (r.decision(lastBase + lastInc[i])) { numCoeff = i + 1; break; }
            }
            i++;
        }
        if (nsig == 0 || where[nsig - 1] != numCoeff - 1) where[nsig++] = (uint8_t)(numCoeff - 1);     // the last position is significant by inference
        PROF_ADD(8, tc);
        int eq1 = 0, gt1 = 0;
        const int gtMax = 4 - (cat == CAT_CDC ? 1 : 0);
        for (int k = nsig - 1; k >= 0; k--) {
            int ctx = absBase + (gt1 != 0 ? 0 : std::min(4, 1 + eq1));
            int v = 0;
            if (r.decision(ctx)) {
                ctx = absBase + 5 + std::min(gtMax, gt1);
                v = 1;
                while (v < 14 && r.decision(ctx)) v++;
                if (v >= 14) {
                    int e = 0; bool cut = false;
                    while (r.bypass()) { v += 1 << e; e++; if (e >= 18) { cut = true; break; } }      // REF: H264Cabac.cpp:4905 stops without the suffix bits
                    if (!cut) while (e--) v += r.bypass() << e;
                }
            }
            const int a = v + 1;
            if (a == 1) eq1++; else gt1++;
            lvl[where[k]] = r.bypass() ? -a : a;
        }
        }
        const int total = nsig;
        if (cat == CAT_I16DC || cat == CAT_CDC) m.cbf_dc ^= (uint8_t)(1 << (comp + 1));
        else m.cbf_ac[comp + 1] ^= (uint16_t)(1 << blk);
        return total;
    }

    // ------------------------------------------------------------ CAVLC residual block (H264ResidualBlockCavlc.cpp:35-232)
    int cavlc_nC(int cat, int blk, int comp) {
        if (cat == CAT_CDC) return -1;
        Nb a, b;
        if (cat == CAT_CAC) { const int x = (blk % 2) * 4, y = (blk / 2) * 4; a = nbr(cur, x - 1, y, true); b = nbr(cur, x, y - 1, true); }
        else { const int bb = cat == CAT_I16DC ? 0 : blk; a = nbr(cur, kBlkX[bb] - 1, kBlkY[bb], false); b = nbr(cur, kBlkX[bb], kBlkY[bb] - 1, false); }
        // REF: the P_Skip / B_Skip / I_PCM special cases compare m_mb_type_fixed (a raw mb_type) with enum values and never
        // match (H264ResidualBlockCavlc.cpp:494-500): the stored TotalCoeff is used for every available neighbour (0 for skipped MBs).
        auto cnt = [&](const Nb &n) {
            const MbT &m = mbs[n.mb];
            if (cat == CAT_CAC) return (int)m.nnz_c[comp][2 * (n.yW / 4) + n.xW / 4];
            return (int)m.nnz[blk_of_xy(n.xW, n.yW)];
        };
        const bool avA = a.mb >= 0, avB = b.mb >= 0;
        if (avA && avB) return (cnt(a) + cnt(b) + 1) >> 1;
        if (avA) return cnt(a);
        if (avB) return cnt(b);
        return 0;
    }
    int residual_block_cavlc(int32_t *lvl, int startIdx, int endIdx, int maxNumCoeff, int cat, int blk, int comp, int *totalCoeff) {
        memset(lvl, 0, sizeof(int32_t) * maxNumCoeff);
        *totalCoeff = 0;
        const int nC = cavlc_nC(cat, blk, comp);
        const int cls = nC < 0 ? 4 : nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3;
        const uint32_t bits = br.peek(16);
        const uint16_t (*tab)[4][2] = coeff_token_table()[cls];
        int tc = -1, t1 = 0;
        for (int c = 0; c <= (cls == 4 ? 4 : 16) && tc < 0; c++)
            for (int t = 0; t < 4; t++) { const int l = tab[c][t][0]; if (l && (bits >> (16 - l)) == tab[c][t][1]) { tc = c; t1 = t; br.skip(l); break; } }
        if (tc < 0) return -1;
        *totalCoeff = tc;
        if (tc == 0) return 0;
        int levelVal[16], runVal[16];
        int suffixLength = (tc > 10 && t1 < 3) ? 1 : 0;
        for (int i = 0; i < tc; i++) {
            if (i < t1) { levelVal[i] = 1 - 2 * (int)br.u1(); continue; }
            int prefix = 0;
            while (!br.u1()) { prefix++; if (prefix > 32) return -1; }
            int sufSize = (prefix == 14 && suffixLength == 0) ? 4 : prefix >= 15 ? prefix - 3 : suffixLength;
            int code = std::min(15, prefix) << suffixLength;
            if (suffixLength > 0 || prefix >= 14) code += sufSize > 0 ? (int)br.u(sufSize) : 0;
            if (prefix >= 15 && suffixLength == 0) code += 15;
            if (prefix >= 16) code += (1 << (prefix - 3)) - 4096;
            if (i == t1 && t1 < 3) code += 2;
            levelVal[i] = (code % 2 == 0) ? (code + 2) >> 1 : (-code - 1) >> 1;
            if (suffixLength == 0) suffixLength = 1;
            if (abs(levelVal[i]) > (3 << (suffixLength - 1)) && suffixLength < 6) suffixLength++;
        }
        int zerosLeft = 0;
        if (tc < endIdx - startIdx + 1) {
            const int kind = maxNumCoeff == 4 ? 1 : maxNumCoeff == 8 ? 2 : 0;
            const uint16_t (*tz)[2] = total_zeros_table()[kind][tc];
            const uint32_t b9 = br.peek(9);
            int found = -1;
            for (int z = 0; z < 16; z++) { const int l = tz[z][0]; if (l && (b9 >> (9 - l)) == tz[z][1]) { found = z; br.skip(l); break; } }
            if (found < 0) return -1;
            zerosLeft = found;
        }
        for (int i = 0; i < tc - 1; i++) {
            if (zerosLeft > 0) {
                const uint16_t (*rb)[2] = run_before_table()[std::min(zerosLeft, 7)];
                const uint32_t b11 = br.peek(11);
                int found = -1;
                for (int z = 0; z < 15; z++) { const int l = rb[z][0]; if (l && (b11 >> (11 - l)) == rb[z][1]) { found = z; br.skip(l); break; } }
                // REF: an unmatched pattern (zerosLeft > 6, eleven zero bits) consumes nothing and leaves the previous run_before of
                // the same CH264ResidualBlockCavlc object in place (H264ResidualBlockCavlc.cpp:3173-3239) — reachable on truncated slices
                if (found < 0) found = last_run_before;
                last_run_before = found;
                runVal[i] = found;
            } else runVal[i] = 0;
            zerosLeft -= runVal[i];
        }
        runVal[tc - 1] = zerosLeft;
        int coeffNum = -1;
        for (int i = tc - 1; i >= 0; i--) {
            coeffNum += runVal[i] + 1;
            if (startIdx + coeffNum >= 0 && startIdx + coeffNum < maxNumCoeff) lvl[startIdx + coeffNum] = levelVal[i];     // (the reference writes out of bounds otherwise)
        }
        return 0;
    }
    int residual_block(int32_t *lvl, int startIdx, int endIdx, int maxNumCoeff, int cat, int blk, int comp, int *total) {
        if (cabac) { *total = residual_block_cabac(lvl, startIdx, endIdx, maxNumCoeff, cat, blk, comp); return 0; }
        return residual_block_cavlc(lvl, startIdx, endIdx, maxNumCoeff, cat, blk, comp, total);
    }

    // ------------------------------------------------------------ residual() (MB:1677-1856; bookkeeping quirks Q18 kept)
    int last_run_before = 0;
    uint32_t coded = 0;          // H264B2_CM_*-style mask of blocks that were parsed with TotalCoeff > 0
    int residual() {
        MbT &m = mbs[cur];
        int total = 0;
        last_run_before = 0;
        // residual_luma works on scratch arrays and the reference copies them to the macroblock only on success (MB:1693-1697):
        // a failing macroblock keeps all-zero levels — emit_mb() is told through residual_ok.  `coded` marks the blocks that were parsed.
        int32_t (&t_dc)[16] = i16dc; int32_t (&t_ac)[16][16] = i16ac; int32_t (&t_l4)[16][16] = l4; int32_t (&t_l8)[4][64] = l8;
        coded = 0;
        if (m.type == T_I16) {
            if (residual_block(t_dc, 0, 15, 16, CAT_I16DC, 0, -1, &total)) return -1;
            m.nnz[0] = (uint8_t)total;
            if (total) coded |= 1u << 16;
        }
        for (int i8 = 0; i8 < 4; i8++) {
            if (!m.t8x8 || !cabac) {
                for (int i4 = 0; i4 < 4; i4++) {
                    const int b = i8 * 4 + i4;
                    if (m.cbp_luma & (1 << i8)) {
                        if (m.type == T_I16) { if (residual_block(t_ac[b], 0, 14, 15, CAT_I16AC, b, -1, &total)) return -1; }
                        else if (residual_block(t_l4[b], 0, 15, 16, CAT_LUMA4, b, -1, &total)) return -1;
                        m.nnz[b] = (uint8_t)total;
                        m.nnz8[i8] = (uint8_t)(m.nnz8[i8] + m.nnz[b]);
                        if (total) coded |= 1u << (m.t8x8 ? i8 : b);
                    }
                    if (!cabac && m.t8x8) {
                        if (m.cbp_luma & (1 << i8)) for (int i = 0; i < 16; i++) t_l8[i8][4 * i + i4] = t_l4[b][i];
                        m.nnz8[i8] = (uint8_t)(m.nnz8[i8] + m.nnz[b]);
                    }
                }
            } else if (m.cbp_luma & (1 << i8)) {
                total = residual_block_cabac(t_l8[i8], 0, 63, 64, CAT_LUMA8, i8, -1);
                m.nnz8[i8] = (uint8_t)total;
                if (total) coded |= 1u << i8;
            }
        }
        last_run_before = 0;        // the chroma blocks use another decoder object (MB:1685)
        for (int c = 0; c < 2; c++) {
            if (m.cbp_chroma & 3) {
                if (residual_block(cdc[c], 0, 3, 4, CAT_CDC, 0, c, &total)) return -1;
                m.nnz_c[c][0] = (uint8_t)total;
                if (total) coded |= 1u << 17;
            } else memset(cdc[c], 0, sizeof cdc[c]);
        }
        for (int c = 0; c < 2; c++)
            for (int b = 0; b < 4; b++) {
                if (m.cbp_chroma & 2) {
                    if (residual_block(cac[c][b], 0, 14, 15, CAT_CAC, b, c, &total)) return -1;
                    m.nnz_c[c][b] = (uint8_t)total;
                    if (total) coded |= 1u << (18 + 4 * c + b);
                }
            }
        return 0;
    }

    // ------------------------------------------------------------ macroblock syntax
    void set_common(MbT &m) { m.field = (uint8_t)F.mb_field; m.slice = (uint16_t)F.slice_number; m.skip_flag = (uint8_t)F.mb_skip_flag; m.decoded = 1; }

    void classify_inter(MbT &m, int st, int t) {     // mb_type of a P or B slice -> type, partition geometry, syntax-level pred modes
        memset(m.part_pm, 0, 4);
        if (st == SLICE_P || st == SLICE_SP) {
            static const uint8_t ty[6] = {T_P16x16, T_P16x8, T_P8x16, T_P8x8, T_P8x8ref0, T_PSKIP};
            m.type = ty[t];
            m.num_part = (t == 0 || t == 5) ? 1 : (t <= 2 ? 2 : 4);
            m.part_w = (t == 0 || t == 1 || t == 5) ? 16 : 8; m.part_h = (t == 0 || t == 2 || t == 5) ? 16 : 8;
            if (t <= 2 || t == 5) memset(m.part_pm, PM_L0, 4);
            m.pm0_inter = (t <= 2 || t == 5);
        } else {
            if (t == 0 || t == 23) { m.type = t == 0 ? T_BDIRECT : T_BSKIP; m.num_part = 0; m.part_w = m.part_h = 8; memset(m.part_pm, PM_DIRECT, 4); m.pm0_inter = 0; }
            else if (t == 22) { m.type = T_B8x8; m.num_part = 4; m.part_w = m.part_h = 8; m.pm0_inter = 0; }
            else if (t <= 3) { m.type = T_B16x16; m.num_part = 1; m.part_w = m.part_h = 16; memset(m.part_pm, t, 4); m.pm0_inter = 1; }
            else {
                static const uint8_t pr[9][2] = {{PM_L0, PM_L0}, {PM_L1, PM_L1}, {PM_L0, PM_L1}, {PM_L1, PM_L0}, {PM_L0, PM_BI}, {PM_L1, PM_BI}, {PM_BI, PM_L0}, {PM_BI, PM_L1}, {PM_BI, PM_BI}};
                const int k = (t - 4) / 2; const bool h16x8 = (t % 2) == 0;
                m.type = h16x8 ? T_B16x8 : T_B8x16; m.num_part = 2; m.part_w = h16x8 ? 16 : 8; m.part_h = h16x8 ? 8 : 16;
                if (h16x8) { m.part_pm[0] = m.part_pm[1] = pr[k][0]; m.part_pm[2] = m.part_pm[3] = pr[k][1]; }
                else { m.part_pm[0] = m.part_pm[2] = pr[k][0]; m.part_pm[1] = m.part_pm[3] = pr[k][1]; }
                m.pm0_inter = 1;
            }
        }
        m.cls = H264B2_MB_INTER; m.intra = 0;
    }
    void classify_intra(MbT &m, int t) {       // I-slice mb_type 0..25
        memset(m.part_pm, 0, 4); m.num_part = 0; m.pm0_inter = 0;
        if (t == 0) { m.type = T_I_NxN; m.intra = 1; m.cls = m.t8x8 ? H264B2_MB_I8x8 : H264B2_MB_I4x4; }
        else if (t == 25) { m.type = T_IPCM; m.intra = 0; m.ipcm = 1; m.cls = H264B2_MB_IPCM; m.i16mode = 3; }    // REF Q11: Intra_NA is not "intra"; Intra16x16PredMode = NA (-1) & 3
        else { m.type = T_I16; m.intra = 1; m.cls = H264B2_MB_I16x16; m.i16mode = (uint8_t)((t - 1) % 4); m.cbp_chroma = (uint8_t)(((t - 1) / 4) % 3); m.cbp_luma = (t - 1) >= 12 ? 15 : 0; }
    }
    // geometry of partition p / sub-partition s of the current macroblock
    void part_rect(const MbT &m, int p, int s, int *x, int *y, int *w, int *h) const {
        const int two = m.part_w != 16;                 // partitions per row - 1 (part_w is 8 or 16): no division
        const int px = (p & two) * m.part_w, py = (p >> two) * m.part_h;
        if (m.type == T_P8x8 || m.type == T_P8x8ref0 || m.type == T_B8x8) {
            static const uint8_t sw[4] = {8, 8, 4, 4}, shh[4] = {8, 4, 8, 4};
            const int w_ = sw[m.sub_shape[p]], h_ = shh[m.sub_shape[p]];
            *x = px + (s % (8 / w_)) * w_; *y = py + (s / (8 / w_)) * h_; *w = w_; *h = h_;
        } else { *x = px; *y = py; *w = m.part_w; *h = m.part_h; }
    }
    int read_ref_idx(int list, int x, int y) {
        if (cabac) return cabac_ref_idx(list, x, y);
        int range = S.listlen[list] - 1;
        if (F.mb_field == 1) range = S.listlen[list] * 2 - 1;      // REF: MB:1321-1326
        return (int)br.te(range);
    }
    int read_mvd(int list, int comp, int x, int y) { return cabac ? cabac_mvd(list, comp, x, y) : br.se(); }

    int mb_pred_inter(MbT &m) {
        const int np = m.num_part;
        const bool fieldDiffers = F.mb_field != sh.field_pic_flag;
        for (int list = 0; list < 2; list++) {
            const int nact = list ? sh.num_ref_idx_l1_active_minus1 : sh.num_ref_idx_l0_active_minus1;
            for (int p = 0; p < np; p++) {
                int x, y, w, h; part_rect(m, p, 0, &x, &y, &w, &h);
                const int pm = m.part_pm[(y / 8) * 2 + x / 8];
                if ((nact > 0 || fieldDiffers) && (pm & (1 << list))) {
                    const int v = read_ref_idx(list, x, y);
                    if (v < 0) return -1;
                    for (int qy = y / 8; qy < (y + h + 7) / 8; qy++) for (int qx = x / 8; qx < (x + w + 7) / 8; qx++) m.ref_syn[list][qy * 2 + qx] = (int8_t)v;
                }
            }
        }
        for (int list = 0; list < 2; list++)
            for (int p = 0; p < np; p++) {
                int x, y, w, h; part_rect(m, p, 0, &x, &y, &w, &h);
                const int pm = m.part_pm[(y / 8) * 2 + x / 8];
                if (pm & (1 << list)) {
                    const int dx = read_mvd(list, 0, x, y);
                    if (dx == MVD_ERROR) return -1;
                    const int dy = read_mvd(list, 1, x, y);
                    if (dy == MVD_ERROR) return -1;
                    for (int yy = y; yy < y + h; yy += 4) for (int xx = x; xx < x + w; xx += 4) { m.mvd[list][(yy / 4) * 4 + xx / 4][0] = (int16_t)dx; m.mvd[list][(yy / 4) * 4 + xx / 4][1] = (int16_t)dy; }
                }
            }
        return 0;
    }
    int sub_mb_pred(MbT &m, int *noSub8x8) {
        const bool isB = m.type == T_B8x8;
        int nsub[4];
        for (int p = 0; p < 4; p++) {
            int t;
            if (cabac) t = isB ? cabac_sub_mb_type_b() : cabac_sub_mb_type_p(); else t = (int)br.ue();
            if (!isB) {
                if (t < 0 || t > 3) return -1;
                m.sub_shape[p] = (uint8_t)t; m.sub_direct[p] = 0; m.part_pm[p] = PM_L0;
            } else {
                if (t < 0 || t > 12) return -1;
                static const uint8_t shp[13] = {3, 0, 0, 0, 1, 2, 1, 2, 1, 2, 3, 3, 3}, pmm[13] = {PM_DIRECT, PM_L0, PM_L1, PM_BI, PM_L0, PM_L0, PM_L1, PM_L1, PM_BI, PM_BI, PM_L0, PM_L1, PM_BI};
                m.sub_shape[p] = shp[t]; m.sub_direct[p] = t == 0; m.part_pm[p] = pmm[t];
            }
            static const int ns[4] = {1, 2, 2, 4};
            nsub[p] = ns[m.sub_shape[p]];
        }
        const bool fieldDiffers = F.mb_field != sh.field_pic_flag;
        for (int list = 0; list < 2; list++) {
            const int nact = list ? sh.num_ref_idx_l1_active_minus1 : sh.num_ref_idx_l0_active_minus1;
            for (int p = 0; p < 4; p++)
                if ((nact > 0 || fieldDiffers) && !(list == 0 && m.type == T_P8x8ref0) && !m.sub_direct[p] && (m.part_pm[p] & (1 << list))) {
                    const int v = read_ref_idx(list, (p % 2) * 8, (p / 2) * 8);
                    if (v < 0) return -1;
                    m.ref_syn[list][p] = (int8_t)v;
                }
        }
        for (int list = 0; list < 2; list++)
            for (int p = 0; p < 4; p++)
                if (!m.sub_direct[p] && (m.part_pm[p] & (1 << list)))
                    for (int s = 0; s < nsub[p]; s++) {
                        int x, y, w, h; part_rect(m, p, s, &x, &y, &w, &h);
                        const int dx = read_mvd(list, 0, x, y);
                        if (dx == MVD_ERROR) return -1;
                        const int dy = read_mvd(list, 1, x, y);
                        if (dy == MVD_ERROR) return -1;
                        for (int yy = y; yy < y + h; yy += 4) for (int xx = x; xx < x + w; xx += 4) { m.mvd[list][(yy / 4) * 4 + xx / 4][0] = (int16_t)dx; m.mvd[list][(yy / 4) * 4 + xx / 4][1] = (int16_t)dy; }
                    }
        *noSub8x8 = 1;
        for (int p = 0; p < 4; p++) {
            if (!m.sub_direct[p]) { if (nsub[p] > 1) *noSub8x8 = 0; }
            else if (!sh.sps.direct_8x8_inference_flag) *noSub8x8 = 0;
        }
        // REF: after sub_mb_pred() the reference looks the sub-macroblock types up again and accepts only 0..3 in B macroblocks
        // (H264MacroBlock.cpp:1061-1069): B_8x8 with an 8x4 / 4x8 / 4x4 sub-macroblock FAILS macroblock_layer() here — no
        // coded_block_pattern, no residual, QPY stays 0 — and parsing goes on from this bit position (SD:380-384).  Reproduced.
        if (isB) for (int p = 0; p < 4; p++) if (!m.sub_direct[p] && m.sub_shape[p] != 0) return -1;
        return 0;
    }

    uint8_t prev_flag[16], rem_mode[16];
    int macroblock_layer() {
        MbT &m = mbs[cur];
        set_common(m);
        const int st = sh.slice_type;
        PROF_T(t1);
        int mb_type = cabac ? cabac_mb_type() : (int)br.ue();
        PROF_ADD(9, t1);
        if (mb_type < 0) return -1;
        int it = -1;      // I mb_type
        if (st == SLICE_I) { if (mb_type > 25) return -1; it = mb_type; }
        else if (st == SLICE_P || st == SLICE_SP) { if (mb_type > 30) return -1; if (mb_type >= 5) it = mb_type - 5; }
        else if (st == SLICE_B) { if (mb_type > 48) return -1; if (mb_type >= 23) it = mb_type - 23; }
        else return -1;
        if (it >= 0) classify_intra(m, it); else classify_inter(m, st, mb_type);
        if (m.type == T_IPCM) {
            // REF (CABAC): the reference re-initialises the arithmetic decoder right after the I_PCM bin — 9 bits BEFORE the alignment
            // bits and the samples (H264Cabac.cpp:3233-3238) instead of after them (9.3.1.2) — and does not initialise it again
            // afterwards; done in cabac_intra_mb_type().  The samples are then read from wherever that leaves the bitstream.
            while (!br.aligned()) br.u1();
            for (int i = 0; i < 384; i++) pcm[i] = (int16_t)br.u(8);
        } else {
            int noSub8x8 = 1;
            PROF_T(t2);
            if (m.type == T_P8x8 || m.type == T_P8x8ref0 || m.type == T_B8x8) { if (sub_mb_pred(m, &noSub8x8)) return -1; }
            else {
                if (sh.pps.transform_8x8_mode_flag && m.type == T_I_NxN) {
                    m.t8x8 = (uint8_t)(cabac ? cabac_t8x8_flag() : br.u1());
                    m.cls = m.t8x8 ? H264B2_MB_I8x8 : H264B2_MB_I4x4;
                }
                if (m.intra) {
                    const int nb = m.type == T_I_NxN ? (m.t8x8 ? 4 : 16) : 0;
                    for (int b = 0; b < nb; b++) {
                        prev_flag[b] = (uint8_t)(cabac ? cb.decision(68) : br.u1());
                        if (!prev_flag[b]) { if (cabac) { int v = cb.decision(69); v |= cb.decision(69) << 1; v |= cb.decision(69) << 2; rem_mode[b] = (uint8_t)v; } else rem_mode[b] = (uint8_t)br.u(3); }
                    }
                    m.chroma_pred = (uint8_t)(cabac ? cabac_intra_chroma_pred_mode() : br.ue());
                } else if (m.type != T_BDIRECT) { if (mb_pred_inter(m)) return -1; }
            }
            PROF_ADD(10, t2);
            PROF_T(t3);
            if (m.type != T_I16) {
                int cbp;
                if (cabac) cbp = cabac_cbp();
                else { const uint32_t cn = br.ue(); if (cn > 47) return -1; cbp = me_cbp_table()[cn][(m.type == T_I_NxN) ? 0 : 1]; }
                m.cbp_luma = (uint8_t)(cbp % 16); m.cbp_chroma = (uint8_t)(cbp / 16);
                if (m.cbp_luma > 0 && sh.pps.transform_8x8_mode_flag && m.type != T_I_NxN && noSub8x8 && (m.type != T_BDIRECT || sh.sps.direct_8x8_inference_flag))
                    m.t8x8 = (uint8_t)(cabac ? cabac_t8x8_flag() : br.u1());
            }
            if (m.cbp_luma > 0 || m.cbp_chroma > 0 || m.type == T_I16) {
                int d = cabac ? cabac_mb_qp_delta() : br.se();
                PROF_ADD(11, t3);
                if (d == QP_DELTA_ERROR) return -1;
                m.qp_delta = (int8_t)std::max(-128, std::min(127, d));
                PROF_T(tr);
                const int rr = residual();
                PROF_ADD(1, tr);
                if (rr) return -1;
                residual_ok = true;
                if (d < -26 || d > 25) d = d < -26 ? -26 : 25;
                m.qp_delta = (int8_t)d;
            }
        }
        const int d = m.qp_delta;
        m.qp = (int8_t)(((F.qp_prev + d + 52) % 52));
        F.qp_prev = m.qp;
        return 0;
    }
    void macroblock_skip() {      // MB:1137-1186
        MbT &m = mbs[cur];
        set_common(m);
        classify_inter(m, sh.slice_type, (sh.slice_type == SLICE_B) ? 23 : 5);
        m.qp_delta = 0;
        m.qp = (int8_t)((F.qp_prev + 52) % 52);
        F.qp_prev = m.qp;
    }

    // ------------------------------------------------------------ intra prediction modes (PB:773-1060)
    void derive_intra_modes() {
        MbT &m = mbs[cur];
        const int cip = sh.pps.constrained_intra_pred_flag;
        const bool i8 = m.t8x8 != 0;
        const int nb = i8 ? 4 : 16;
        for (int b = 0; b < nb; b++) {
            const int x = i8 ? (b % 2) * 8 : kBlkX[b], y = i8 ? (b / 2) * 8 : kBlkY[b];
            const Nb A = nbr(cur, x - 1, y, false), B = nbr(cur, x, y - 1, false);
            bool dc = A.mb < 0 || B.mb < 0 || (A.mb >= 0 && mbs[A.mb].pm0_inter && cip) || (B.mb >= 0 && mbs[B.mb].pm0_inter && cip);
            int mode[2];
            for (int k = 0; k < 2; k++) {
                const Nb &N = k ? B : A;
                if (dc || mbs[N.mb].type != T_I_NxN) { mode[k] = 2; continue; }
                const MbT &n = mbs[N.mb];
                const int b4 = blk_of_xy(N.xW, N.yW), b8 = (N.yW / 8) * 2 + N.xW / 8;
                if (!i8) mode[k] = n.t8x8 ? n.ipred[b4 >> 2] : n.ipred[b4];
                else if (n.t8x8) mode[k] = n.ipred[b8];
                else mode[k] = n.ipred[b8 * 4 + (k ? 2 : 1)];      // REF Q17: n = 1 for A in every case (PB:1004 tests field_pic_flag)
            }
            const int pred = std::min(mode[0], mode[1]);
            m.ipred[b] = (int8_t)(prev_flag[b] ? pred : (rem_mode[b] < pred ? rem_mode[b] : rem_mode[b] + 1));
        }
    }

    // ------------------------------------------------------------ motion (IP:412-2047)
    struct NbMv { int avail; int ref; int mv[2]; };
    NbMv fetch(int x, int y, int list) const {
        NbMv r; r.avail = 0; r.ref = -1; r.mv[0] = r.mv[1] = 0;
        const Nb n = nbr(cur, x, y, false);
        if (n.mb < 0) return r;
        r.avail = 1;
        const MbT &m = mbs[n.mb];
        const int q = (n.yW / 8) * 2 + n.xW / 8, b = (n.yW / 4) * 4 + n.xW / 4;
        if (m.intra || !m.pf[list][q]) return r;
        r.ref = m.ref[list][q]; r.mv[0] = mot[n.mb].mv[list][b][0]; r.mv[1] = mot[n.mb].mv[list][b][1];
        if (mbs[cur].field == 1 && m.field == 0) { r.mv[1] = r.mv[1] / 2; r.ref = r.ref * 2; }
        else if (mbs[cur].field == 0 && m.field == 1) { r.mv[1] = r.mv[1] * 2; r.ref = r.ref / 2; }
        return r;
    }
    // neighbouring partitions A, B, C (C falls back to D) of the rectangle at (x, y) with width predPartWidth (IP:1764-1990).
    // REF: partitions of the CURRENT macroblock that are not decoded yet are "available" with predFlag 0 (the check at
    // IP:3137-3140 is empty), so no D substitution happens for them.
    void neighbours(int x, int y, int predPartWidth, int list, NbMv &A, NbMv &B, NbMv &C) const {
        A = fetch(x - 1, y, list); B = fetch(x, y - 1, list); C = fetch(x + predPartWidth, y - 1, list);
        if (!C.avail) C = fetch(x - 1, y - 1, list);
    }
    void predict_mv(const MbT &m, int p, int x, int y, int predPartWidth, int list, int refIdx, int mvp[2]) const {
        NbMv A, B, C; neighbours(x, y, predPartWidth, list, A, B, C);
        if (m.part_w == 16 && m.part_h == 8 && p == 0 && B.ref == refIdx) { mvp[0] = B.mv[0]; mvp[1] = B.mv[1]; return; }
        if (m.part_w == 16 && m.part_h == 8 && p == 1 && A.ref == refIdx) { mvp[0] = A.mv[0]; mvp[1] = A.mv[1]; return; }
        if (m.part_w == 8 && m.part_h == 16 && p == 0 && A.ref == refIdx) { mvp[0] = A.mv[0]; mvp[1] = A.mv[1]; return; }
        if (m.part_w == 8 && m.part_h == 16 && p == 1 && C.ref == refIdx) { mvp[0] = C.mv[0]; mvp[1] = C.mv[1]; return; }
        if (!B.avail && !C.avail && A.avail) { B = A; C = A; }
        const bool a = A.ref == refIdx, b = B.ref == refIdx, c = C.ref == refIdx;
        if (a && !b && !c) { mvp[0] = A.mv[0]; mvp[1] = A.mv[1]; }
        else if (!a && b && !c) { mvp[0] = B.mv[0]; mvp[1] = B.mv[1]; }
        else if (!a && !b && c) { mvp[0] = C.mv[0]; mvp[1] = C.mv[1]; }
        else for (int k = 0; k < 2; k++) mvp[k] = A.mv[k] + B.mv[k] + C.mv[k] - std::min(A.mv[k], std::min(B.mv[k], C.mv[k])) - std::max(A.mv[k], std::max(B.mv[k], C.mv[k]));
    }
    // Reference_picture_selection_process (IP:2117-2197): slot + view of RefPicListX[refIdx]; -1 on failure
    int list_ok[2] = {-1, -1};        // per slice: every entry of RefPicListX[0..len) is a marked frame (IP:2134-2147 checks this on every call)
    int select_ref(int list, int refIdx) {
        if (refIdx < 0 || refIdx >= 32) return -1;
        if (list_ok[list] < 0) {
            list_ok[list] = 1;
            const int len = S.listlen[list];
            for (int i = 0; i < len; i++) { const int s = S.list[list][i]; if (s < 0 || F.slots[s].p_coded_marked != 1) { list_ok[list] = 0; break; } }
        }
        if (!list_ok[list]) return -1;
        if (!mbs[cur].field) { const int s = refIdx < 34 ? S.list[list][refIdx] : -1; return s < 0 ? -1 : (s << 2); }
        const int s = S.list[list][refIdx / 2];
        if (s < 0) return -1;
        const int same = (refIdx % 2) == 0, bottomMb = cur % 2;
        const int view = (same ? bottomMb : !bottomMb) ? 2 : 1;
        return (s << 2) | view;
    }
    // co-located 4x4 (IP:1017-1296), frame / MBAFF pictures only
    int colocated(int p, int s, int *mbAddrCol, int mvCol[2], int *refIdxCol) const {
        const int l10 = S.list[1][0];
        if (l10 < 0) return -1;
        const Slot &R = F.slots[l10];
        int topAbs = 0, botAbs = 0;
        if (R.p_coded_marked == 1) { topAbs = abs(R.TopFieldOrderCnt - S.PicOrderCnt); botAbs = abs(R.BottomFieldOrderCnt - S.PicOrderCnt); }
        if (!(R.p_finished == 1 && R.p_coded_marked == 1)) return -1;
        if ((int)R.col.size() != nmb) return -1;
        const bool curA = sh.sps.mb_adaptive_frame_field_flag != 0, colA = R.hdr_mbaff_sps != 0;
        if (curA != colA) return -1;
        const int blk = sh.sps.direct_8x8_inference_flag ? 5 * p : 4 * p + s;
        const int xCol = kBlkX[blk], yCol = kBlkY[blk];
        int addr = cur, yM = yCol;
        if (curA) {
            const int colField = R.col[cur].field;
            if (F.mb_field == 0) { if (colField) { addr = 2 * (cur / 2) + (topAbs < botAbs ? 0 : 1); yM = 8 * (cur % 2) + 4 * (yCol / 8); } }
            else if (!colField) { addr = 2 * (cur / 2) + (yCol / 8); yM = (2 * yCol) % 16; }
        }
        const ColMb &c = R.col[addr];
        if (c.type == T_NA) return -1;
        *mbAddrCol = addr;
        if (c.intra) { mvCol[0] = mvCol[1] = 0; *refIdxCol = -1; return 0; }
        const int q = (yM / 8) * 2 + xCol / 8, b = (yM / 4) * 4 + xCol / 4;
        const int l = ((c.pf0 >> q) & 1) ? 0 : 1;
        mvCol[0] = R.motion[addr].mv[l][b][0]; mvCol[1] = R.motion[addr].mv[l][b][1]; *refIdxCol = c.ref[l][q];
        return 0;
    }
    std::vector<int32_t> wcache;       // per slice: (refIdxL0, refIdxL1, field MB parity) -> weight table index
    int weight_index(int ref0, int ref1, int pf0, int pf1) {
        if (ref0 < -1 || ref0 > 31 || ref1 < -1 || ref1 > 31) return weight_index_slow(ref0, ref1, pf0, pf1);
        if (wcache.empty()) wcache.assign(33 * 33 * 3, -1);
        const int par = mbs[cur].field ? 1 + (cur & 1) : 0;
        int32_t &c = wcache[((ref0 + 1) * 33 + (ref1 + 1)) * 3 + par];
        if (c < 0) { const int v = weight_index_slow(ref0, ref1, pf0, pf1); if (v < 0) return v; c = v; }
        return c;
    }
    int weight_index_slow(int ref0, int ref1, int pf0, int pf1) {
        // IP:538-546 + Derivation_process_for_prediction_weights (IP:2833-3047) + the mode selection of IP:2545-2610
        const int st = sh.slice_type % 5;
        int logWD[3] = {0, 0, 0}, w0[3] = {1, 1, 1}, w1[3] = {1, 1, 1}, o0[3] = {0, 0, 0}, o1[3] = {0, 0, 0};
        const bool derive = (sh.pps.weighted_pred_flag == 1 && (st == SLICE_P || st == SLICE_SP)) || (sh.pps.weighted_bipred_idc > 0 && st == SLICE_B);
        if (derive) {
            int implicitMode = 0, explicitMode = 0;
            if (sh.pps.weighted_bipred_idc == 2 && st == SLICE_B && pf0 && pf1) implicitMode = 1;
            else if (sh.pps.weighted_bipred_idc == 1 && st == SLICE_B && (pf0 + pf1 >= 1)) explicitMode = 1;
            else if (sh.pps.weighted_pred_flag == 1 && (st == SLICE_P || st == SLICE_SP) && pf0) explicitMode = 1;
            if (implicitMode) {
                int curPoc, poc0, poc1;
                const MbT &m = mbs[cur];
                if (m.field) {
                    const int s0 = S.list[0][ref0 / 2], s1 = S.list[1][ref1 / 2];
                    if (s0 < 0 || s1 < 0) return -1;
                    const bool bot = cur % 2;
                    curPoc = bot ? S.BottomFieldOrderCnt : S.TopFieldOrderCnt;
                    const bool b0 = (ref0 % 2 == 0) ? bot : !bot, b1 = (ref1 % 2 == 0) ? bot : !bot;
                    poc0 = b0 ? F.slots[s0].BottomFieldOrderCnt : F.slots[s0].TopFieldOrderCnt;
                    poc1 = b1 ? F.slots[s1].BottomFieldOrderCnt : F.slots[s1].TopFieldOrderCnt;
                } else {
                    const int s0 = ref0 < 34 ? S.list[0][ref0] : -1, s1 = ref1 < 34 ? S.list[1][ref1] : -1;
                    if (s0 < 0 || s1 < 0) return -1;
                    curPoc = std::min(S.TopFieldOrderCnt, S.BottomFieldOrderCnt);
                    poc0 = std::min(F.slots[s0].TopFieldOrderCnt, F.slots[s0].BottomFieldOrderCnt);
                    poc1 = std::min(F.slots[s1].TopFieldOrderCnt, F.slots[s1].BottomFieldOrderCnt);
                }
                auto clip3 = [](int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; };
                const int tb = clip3(-128, 127, curPoc - poc0), td = clip3(-128, 127, poc1 - poc0);
                int wa = 32, wb = 32;
                if (td != 0) {       // (the reference divides by td before testing for 0 and would trap here)
                    const int tx = (16384 + abs(td / 2)) / td;
                    const int dsf = clip3(-1024, 1023, (tb * tx + 32) >> 6);
                    const int s0 = m.field ? S.list[0][ref0 / 2] : S.list[0][ref0], s1 = m.field ? S.list[1][ref1 / 2] : S.list[1][ref1];
                    const bool lt = F.slots[s0].mark == MARK_LONG || F.slots[s1].mark == MARK_LONG;
                    if (!(poc1 - poc0 == 0 || lt || (dsf >> 2) < -64 || (dsf >> 2) > 128)) { wa = 64 - (dsf >> 2); wb = dsf >> 2; }
                }
                for (int c = 0; c < 3; c++) { logWD[c] = 5; w0[c] = wa; w1[c] = wb; o0[c] = 0; o1[c] = 0; }
            } else if (explicitMode) {
                const bool half = mbaff && mbs[cur].field;
                const int r0 = half ? ref0 >> 1 : ref0, r1 = half ? ref1 >> 1 : ref1;
                // REF Q8: a list-1-only partition takes its LUMA weight/offset from luma_weight_l0[refIdxL0] / luma_offset_l0[refIdxL0] with
                // refIdxL0 == -1 (IP:2764/2768 + IP:3003-3006), i.e. the struct members in front of the arrays (H264SliceHeader.h:78-80):
                // luma_weight_l0[-1] is luma_weight_l0_flag as last parsed, luma_offset_l0[-1] is luma_weight_l0[31]
                auto lw = [&](int l, int r) { return (r >= 0 && r < 32) ? sh.luma_weight[l][r] : (r == -1 ? sh.last_luma_weight_flag[l] : 0); };
                auto lo = [&](int l, int r) { return (r >= 0 && r < 32) ? sh.luma_offset[l][r] : (r == -1 ? sh.luma_weight[l][31] : 0); };
                auto cw = [&](int l, int r, int j) { return (r >= 0 && r < 32) ? sh.chroma_weight[l][r][j] : 0; };
                auto co = [&](int l, int r, int j) { return (r >= 0 && r < 32) ? sh.chroma_offset[l][r][j] : 0; };
                logWD[0] = sh.luma_log2_weight_denom; w0[0] = lw(0, r0); w1[0] = lw(1, r1); o0[0] = lo(0, r0); o1[0] = lo(1, r1);
                for (int j = 0; j < 2; j++) { logWD[1 + j] = sh.chroma_log2_weight_denom; w0[1 + j] = cw(0, r0, j); w1[1 + j] = cw(1, r1, j); o0[1 + j] = co(0, r0, j); o1[1 + j] = co(1, r1, j); }
            }
        }
        int mode = 0;
        if (pf0 == 1 && (st == SLICE_P || st == SLICE_SP)) mode = sh.pps.weighted_pred_flag ? 1 : 0;
        else if ((pf0 || pf1) && st == SLICE_B) {
            if (sh.pps.weighted_bipred_idc == 1) mode = 1;
            else if (sh.pps.weighted_bipred_idc == 2) mode = (pf0 && pf1) ? 1 : 0;
        }
        if (mode == 0) return 0;
        H264B2Weight we; memset(&we, 0, sizeof we);
        we.mode = 1;
        for (int c = 0; c < 3; c++) { we.logwd[c] = (int16_t)logWD[c]; we.w0[c] = (int16_t)w0[c]; we.w1[c] = (int16_t)w1[c]; we.o0[c] = (int16_t)o0[c]; we.o1[c] = (int16_t)o1[c]; }
        if (!pf0 && pf1) { we.w1[0] = (int16_t)w0[0]; we.o1[0] = (int16_t)o0[0]; }       // REF Q8: IP:2764/2768 use w0L/o0L for list-1-only luma
        if (pf0 && !pf1) for (int c = 0; c < 3; c++) { we.w1[c] = 0; we.o1[c] = 0; }
        if (!pf0 && pf1) for (int c = 0; c < 3; c++) { we.w0[c] = 0; we.o0[c] = 0; }
        for (size_t i = 0; i < F.weights.size(); i++) if (!memcmp(&F.weights[i], &we, sizeof we)) return (int)i;
        if (F.weights.size() >= 65535) return -1;
        F.weights.push_back(we);
        return (int)F.weights.size() - 1;
    }

    bool dc_valid = false; int dc_r[2], dc_zero, dc_mvp[2][2];
    int inter_prediction() {       // the derivation half of Inter_prediction_process (IP:412-667)
        MbT &m = mbs[cur];
        dc_valid = false;
        H264B2MbMotion &M = mot[cur];
        memset(M.ref_surf, -1, sizeof M.ref_surf);
        // m_RefIdxLX are 0 until a partition is derived (memset of m_mbs): a macroblock whose derivation stops early still reports the
        // identity of RefPicListX[0] for the partitions it never reached (oracle/ref_harness.cpp reads refIdx >= 0 whatever predFlag is)
        for (int l = 0; l < 2; l++) for (int q = 0; q < 4; q++) M.ref_ident[l][q] = (int8_t)S.list[l][0];
        F.has_inter = 1;
        const bool direct16 = m.type == T_BSKIP || m.type == T_BDIRECT;
        const bool is8x8 = m.type == T_P8x8 || m.type == T_P8x8ref0 || m.type == T_B8x8;
        const int np = direct16 ? 4 : m.num_part;
        int mvL0[2] = {0, 0}, mvL1[2] = {0, 0};        // function-scope in the reference: stale values of earlier partitions are stored for unused lists
        int refIdx[2] = {-1, -1};
        for (int p = 0; p < np; p++) {
            const bool subDirect = m.type == T_B8x8 && m.sub_direct[p];
            int nsub;
            if (!is8x8 && !direct16) nsub = 1;
            else if (is8x8 && !subDirect) { static const int ns[4] = {1, 2, 2, 4}; nsub = ns[m.sub_shape[p]]; }
            else nsub = 4;
            // direct_8x8_inference: the four 4x4 blocks of a direct quadrant share the co-located corner block (5 * p), the reference
            // indices and both predictors, so they get identical results — derived once for the whole 8x8
            const bool direct8 = (direct16 || subDirect) && sh.sps.direct_8x8_inference_flag;
            if (direct8) nsub = 1;
            for (int s = 0; s < nsub; s++) {
                int x, y, w, h;
                if (direct8) { x = (p % 2) * 8; y = (p / 2) * 8; w = h = 8; }
                else if (direct16 || subDirect) { x = (p % 2) * 8 + (s % 2) * 4; y = (p / 2) * 8 + (s / 2) * 4; w = h = 4; }
                else part_rect(m, p, s, &x, &y, &w, &h);
                int pf[2] = {0, 0};
                if (m.type == T_PSKIP) {
                    refIdx[0] = 0;
                    for (int q = 0; q < 4; q++) m.pf[0][q] = 1;
                    NbMv A = fetch(-1, 0, 0), B = fetch(0, -1, 0);
                    if (!A.avail || !B.avail || (A.ref == 0 && A.mv[0] == 0 && A.mv[1] == 0) || (B.ref == 0 && B.mv[0] == 0 && B.mv[1] == 0)) { mvL0[0] = mvL0[1] = 0; }
                    else predict_mv(m, 0, 0, 0, 16, 0, 0, mvL0);
                    pf[0] = 1; pf[1] = 0; mvL1[0] = mvL1[1] = -1;       // REF: mvL1 = NA (-1) is what lands in m_MvL1
                    refIdx[1] = -1;      // stays at its initial value in the reference
                } else if (direct16 || subDirect) {
                    if (!sh.direct_spatial_mv_pred_flag) { F.error = "temporal direct prediction is not supported"; return -2; }
                    // spatial direct (IP:1302-1428): neighbours of the whole macroblock, as partition 0 of width 16.  They lie outside
                    // the macroblock, so the reference indices and the two predictors are the same for all of its direct 4x4 blocks.
                    if (!dc_valid) {
                        for (int l = 0; l < 2; l++) {
                            NbMv A, B, C; neighbours(0, 0, 16, l, A, B, C);
                            auto minpos = [](int a, int b) { return (a >= 0 && b >= 0) ? std::min(a, b) : std::max(a, b); };
                            dc_r[l] = minpos(A.ref, minpos(B.ref, C.ref));
                        }
                        dc_zero = 0;
                        if (dc_r[0] < 0 && dc_r[1] < 0) { dc_r[0] = dc_r[1] = 0; dc_zero = 1; }
                        for (int l = 0; l < 2; l++) { dc_mvp[l][0] = dc_mvp[l][1] = 0; if (!dc_zero && dc_r[l] >= 0) predict_mv(m, 0, 0, 0, 16, l, dc_r[l], dc_mvp[l]); }
                        dc_valid = true;
                    }
                    int r[2] = {dc_r[0], dc_r[1]};
                    const int directZero = dc_zero;
                    int addrCol = 0, mvCol[2] = {0, 0}, refCol = 0;
                    if (colocated(p, s, &addrCol, mvCol, &refCol)) return -1;
                    const int l10 = S.list[1][0];
                    // REF Q10: no frame/field unit conversion of mvCol
                    const int colZero = (F.slots[l10].p_mark == MARK_SHORT && refCol == 0 && mvCol[0] >= -1 && mvCol[0] <= 1 && mvCol[1] >= -1 && mvCol[1] <= 1) ? 1 : 0;
                    if (directZero || r[0] < 0 || (r[0] == 0 && colZero)) mvL0[0] = mvL0[1] = 0; else { mvL0[0] = dc_mvp[0][0]; mvL0[1] = dc_mvp[0][1]; }
                    if (directZero || r[1] < 0 || (r[1] == 0 && colZero)) mvL1[0] = mvL1[1] = 0; else { mvL1[0] = dc_mvp[1][0]; mvL1[1] = dc_mvp[1][1]; }
                    refIdx[0] = r[0]; refIdx[1] = r[1];
                    pf[0] = r[0] >= 0; pf[1] = r[1] >= 0;
                } else {
                    const int q = (y / 8) * 2 + x / 8;
                    const int pm = m.part_pm[q];
                    pf[0] = (pm & 1) != 0; pf[1] = (pm & 2) != 0;
                    refIdx[0] = pf[0] ? m.ref_syn[0][q] : -1; refIdx[1] = pf[1] ? m.ref_syn[1][q] : -1;
                    const int ppw = is8x8 ? w : m.part_w;
                    for (int l = 0; l < 2; l++) {
                        if (!pf[l]) continue;
                        if (select_ref(l, refIdx[l]) < 0) return -1;
                        int mvp[2]; predict_mv(m, p, x, y, ppw, l, refIdx[l], mvp);
                        int *mv = l ? mvL1 : mvL0;
                        mv[0] = mvp[0] + m.mvd[l][(y / 4) * 4 + x / 4][0]; mv[1] = mvp[1] + m.mvd[l][(y / 4) * 4 + x / 4][1];
                    }
                }
                // reference surfaces / bS identities / weights of this partition
                int8_t rs[2] = {-1, -1}, ri[2] = {-1, -1};
                for (int l = 0; l < 2; l++) {
                    if (pf[l]) { const int v = select_ref(l, refIdx[l]); if (v < 0) return -1; rs[l] = (int8_t)v; }
                    if (refIdx[l] >= 0) ri[l] = (int8_t)((refIdx[l] < 16) ? S.list[l][refIdx[l]] : -1);      // REF Q6: raw refIdx, last-built list
                }
                const int widx = weight_index(refIdx[0], refIdx[1], pf[0], pf[1]);
                if (widx < 0) return -1;
                // store (IP:593-600) — flattened to 4x4 blocks / 8x8 quadrants
                const int pw8 = (direct16 || subDirect || is8x8) ? 8 : m.part_w, ph8 = (direct16 || subDirect || is8x8) ? 8 : m.part_h;
                const int px = (direct16 || subDirect || is8x8) ? (p % 2) * 8 : x, py = (direct16 || subDirect || is8x8) ? (p / 2) * 8 : y;
                for (int qy = py / 8; qy < (py + ph8 + 7) / 8; qy++)
                    for (int qx = px / 8; qx < (px + pw8 + 7) / 8; qx++) {
                        const int q = qy * 2 + qx;
                        for (int l = 0; l < 2; l++) { m.pf[l][q] = (uint8_t)pf[l]; m.ref[l][q] = (int8_t)refIdx[l]; M.ref_surf[l][q] = rs[l]; M.ref_ident[l][q] = ri[l]; }
                        M.wt_idx[q] = (uint16_t)widx;
                    }
                for (int yy = y; yy < y + h; yy += 4)
                    for (int xx = x; xx < x + w; xx += 4) {
                        const int b = (yy / 4) * 4 + xx / 4;
                        M.mv[0][b][0] = (int16_t)mvL0[0]; M.mv[0][b][1] = (int16_t)mvL0[1]; M.mv[1][b][0] = (int16_t)mvL1[0]; M.mv[1][b][1] = (int16_t)mvL1[1];
                    }
            }
        }
        return 0;
    }

    // ------------------------------------------------------------ SoA emission of one macroblock
    // Levels travel as int16 (saturated, like the container the reference harness writes) and only blocks with a non-zero level are
    // present.  One pass: saturating pack, OR of everything, store at the write position; the position advances only when something was
    // non-zero.  n = 8 (both chroma DC blocks), 16 or 64; shift_in: AC blocks of 15 levels are stored as [0, level 0..14].
    int16_t *cdst = nullptr;
    inline bool put_if_nz(const int32_t *src, int n, int shift_in) {
        const __m128i zero = _mm_setzero_si128();
        __m128i any;
        if (shift_in) {
            const __m128i a0 = _mm_loadu_si128((const __m128i *)src), a1 = _mm_loadu_si128((const __m128i *)(src + 4)), a2 = _mm_loadu_si128((const __m128i *)(src + 8));
            const __m128i a3 = _mm_and_si128(_mm_loadu_si128((const __m128i *)(src + 12)), _mm_set_epi32(0, -1, -1, -1));      // src[15] is not part of the block
            const __m128i p0 = _mm_packs_epi32(a0, a1), p1 = _mm_packs_epi32(a2, a3);
            any = _mm_or_si128(p0, p1);
            _mm_storeu_si128((__m128i *)cdst, _mm_slli_si128(p0, 2));
            _mm_storeu_si128((__m128i *)(cdst + 8), _mm_or_si128(_mm_slli_si128(p1, 2), _mm_srli_si128(p0, 14)));
        } else {
            any = zero;
            for (int k = 0; k < n; k += 8) {
                const __m128i pk = _mm_packs_epi32(_mm_loadu_si128((const __m128i *)(src + k)), _mm_loadu_si128((const __m128i *)(src + k + 4)));
                any = _mm_or_si128(any, pk);
                _mm_storeu_si128((__m128i *)(cdst + k), pk);
            }
        }
        if (_mm_movemask_epi8(_mm_cmpeq_epi8(any, zero)) == 0xffff) return false;
        cdst += n;
        return true;
    }
    void emit_mb(bool have_residual) {
        const MbT &m = mbs[cur];
        H264B2MbInfo &I = F.info[cur];
        if (F.coefs.size() < F.ncoef + 1024) F.coefs.resize(std::max(F.coefs.size() * 2, F.ncoef + ((size_t)1 << 18)));      // a macroblock writes <= 408 levels
        cdst = F.coefs.data() + F.ncoef;
        F.coff[cur] = (uint32_t)F.ncoef;
        I.mb_class = m.cls;
        const bool spsi = sh.slice_type == SLICE_SP || sh.slice_type == SLICE_SI;
        I.flags = (uint8_t)((m.field ? H264B2_MBF_FIELD : 0) | (m.t8x8 ? H264B2_MBF_T8x8 : 0) | (spsi ? H264B2_MBF_SPSI : 0) | ((m.pm0_inter && sh.pps.constrained_intra_pred_flag) ? H264B2_MBF_CIP_UNAVAIL : 0));
        I.pred16_chroma = (uint8_t)((m.i16mode & 3) | ((m.chroma_pred & 3) << 2));
        I.qpy = m.qp;
        I.slice_number = m.slice;
        uint16_t nnz = 0;
        if (m.t8x8) { for (int q = 0; q < 4; q++) if (m.nnz8[q]) nnz |= (uint16_t)(0xFu << (4 * q)); }
        else nnz = (uint16_t)~_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i *)m.nnz), _mm_setzero_si128()));
        I.nnz_mask = nnz;
        I.filter_offset_a = (int8_t)sh.FilterOffsetA; I.filter_offset_b = (int8_t)sh.FilterOffsetB; I.deblock_idc = (uint8_t)sh.disable_deblocking_filter_idc;
        uint32_t cm = 0;
        if (m.cls == H264B2_MB_IPCM) { cm |= H264B2_CM_PCM; memcpy(cdst, pcm, sizeof pcm); cdst += 384; }
        else if (have_residual && coded) {
            // a block is present iff it has a non-zero level (oracle/ref_harness.cpp); only parsed blocks (TotalCoeff > 0) can have one
            const uint32_t cd = coded;
            if (m.cls == H264B2_MB_I16x16) {
                for (int b = 0; b < 16; b++) if ((cd >> b) & 1) if (put_if_nz(i16ac[b], 16, 1)) cm |= H264B2_CM_LUMA(b);
                if ((cd >> 16) & 1) if (put_if_nz(i16dc, 16, 0)) cm |= H264B2_CM_LUMA_DC;
            } else if (m.t8x8) { for (int b = 0; b < 4; b++) if ((cd >> b) & 1) if (put_if_nz(l8[b], 64, 0)) cm |= H264B2_CM_LUMA(b); }
            else for (int b = 0; b < 16; b++) if ((cd >> b) & 1) if (put_if_nz(l4[b], 16, 0)) cm |= H264B2_CM_LUMA(b);
            if ((cd >> 17) & 1) if (put_if_nz(cdc[0], 8, 0)) cm |= H264B2_CM_CHROMA_DC;      // cdc[0] and cdc[1] are contiguous: Cb DC then Cr DC
            for (int b = 0; b < 4; b++) if ((cd >> (18 + b)) & 1) if (put_if_nz(cac[0][b], 16, 1)) cm |= H264B2_CM_CB(b);
            for (int b = 0; b < 4; b++) if ((cd >> (22 + b)) & 1) if (put_if_nz(cac[1][b], 16, 1)) cm |= H264B2_CM_CR(b);
        }
        F.ncoef = (size_t)(cdst - F.coefs.data());
        I.coef_mask = cm;
        if (m.cls == H264B2_MB_I4x4) { uint64_t v = 0; for (int b = 0; b < 16; b++) v |= (uint64_t)(m.ipred[b] & 15) << (4 * b); F.modes[cur] = v; }
        if (m.cls == H264B2_MB_I8x8) { uint64_t v = 0; for (int b = 0; b < 4; b++) v |= (uint64_t)(m.ipred[b] & 15) << (4 * b); F.modes[cur] = v; }
    }

    // ------------------------------------------------------------ slice_data (SD:64-534)
    int infer_field_flag() const {      // 7.4.4 inference when neither macroblock of a pair carries the flag
        if ((cur / 2) % W > 0 && mbs[cur - 2].slice == F.slice_number) return mbs[cur - 2].field;
        if (cur / (2 * W) > 0 && mbs[cur - 2 * W].slice == F.slice_number) return mbs[cur - 2 * W].field;
        return 0;
    }
    int next_mb(int n) const { const int i = n + 1; return i >= nmb ? -2 : i; }

    int skip_mb() {
        S.mb_cnt++;
        macroblock_skip();
        const int r = inter_prediction();
        if (r == 0) emit_mb(false);
        else emit_mb(false);
        return r;
    }

    int run() {
        const int st = sh.slice_type;
        F.slice_number = ++S.slice_number;
        if (cabac) {
            if ((unsigned)sh.cabac_init_idc > 2) return -1;      // non-conforming (the reference goes on with uninitialised contexts, H264Cabac.cpp:41)
            while (!br.aligned()) br.u1();
            cb.init_contexts(st, sh.cabac_init_idc, sh.SliceQPY); cb.start_slice(&br);
        }
        if (!mbaff) F.mb_field = sh.field_pic_flag;
        cur = sh.first_mb_in_slice * (1 + mbaff);
        F.qp_prev = sh.SliceQPY;
        int moreData = 1, prevMbSkipped = 0, skipNext = 0;
        if (S.slice_cnt == 0) {
            if (F.decode_poc()) return -1;
            if (st == SLICE_P || st == SLICE_SP || st == SLICE_B) if (F.build_ref_lists()) return -1;
        }
        S.slice_cnt++;
        if (cur >= nmb) return -1;
        bool skipReadField = false;
        do {
            if (st != SLICE_I && st != SLICE_SI) {
                if (!cabac) {
                    const int run = (int)br.ue();
                    prevMbSkipped = run > 0;
                    for (int i = 0; i < run; i++) {
                        if (mbaff && cur % 2 == 0) {
                            if (i == run - 1) { F.mb_field = br.u1(); skipReadField = true; }
                            else F.mb_field = infer_field_flag();
                        }
                        F.mb_skip_flag = 0;
                        if (skip_mb()) return -1;
                        cur = next_mb(cur);
                        if (cur < 0) break;
                    }
                    if (run > 0) moreData = br.more_rbsp_data();
                    if (cur < 0) { if (!moreData) break; return -1; }
                } else {
                    mbs[cur].slice = (uint16_t)F.slice_number;
                    if (mbaff) {
                        if (cur % 2 == 0 && (cur / 2) % W == 0 && cur / (2 * W) >= 1) F.mb_field = infer_field_flag();     // REF: SD:250-266, only at the start of a pair row
                        mbs[cur].field = (uint8_t)F.mb_field;
                    }
                    if (mbaff && cur % 2 == 1 && prevMbSkipped) F.mb_skip_flag = skipNext;
                    else { PROF_T(t4); F.mb_skip_flag = cabac_mb_skip_flag(cur); PROF_ADD(12, t4); }
                    if (F.mb_skip_flag) {
                        if (mbaff && cur % 2 == 0) {
                            mbs[cur].skip_flag = 1;
                            mbs[cur + 1].slice = (uint16_t)F.slice_number; mbs[cur + 1].field = (uint8_t)F.mb_field;
                            skipNext = cabac_mb_skip_flag(cur + 1);
                            if (!skipNext) { F.mb_field = cabac_mb_field_flag(); skipReadField = true; }
                            else F.mb_field = infer_field_flag();
                        }
                        PROF_T(ts);
                        const int sr = skip_mb();
                        PROF_ADD(5, ts);
                        if (sr) return -1;
                    }
                    moreData = !F.mb_skip_flag;
                }
            }
            if (moreData) {
                if (mbaff && (cur % 2 == 0 || (cur % 2 == 1 && prevMbSkipped))) {
                    if (!skipReadField) F.mb_field = cabac ? cabac_mb_field_flag_pre() : (int)br.u1();
                    else skipReadField = false;
                }
                S.mb_cnt++;
                residual_ok = false;
                mbs[cur].slice = (uint16_t)F.slice_number; mbs[cur].field = (uint8_t)F.mb_field;
                PROF_T(tm);
                const int r = macroblock_layer();
                PROF_ADD(0, tm);       // a failure is logged by the reference and reconstruction goes on with what was parsed (SD:380-384)
                if (r != 0) { MbT &m = mbs[cur]; if (m.type == T_NA) { /* nothing usable was parsed: the reference's reconstruction calls fail on MB_TYPE_NA */ return -1; } }
                MbT &m = mbs[cur];
                if (m.cls == H264B2_MB_INTER) { PROF_T(ti); const int e = inter_prediction(); PROF_ADD(2, ti); PROF_T(te); emit_mb(residual_ok); PROF_ADD(4, te); if (e) return -1; }
                else { PROF_T(ti); if (m.type == T_I_NxN) derive_intra_modes(); PROF_ADD(3, ti); PROF_T(te); emit_mb(residual_ok); PROF_ADD(4, te); }
            }
            if (!cabac) moreData = br.more_rbsp_data();
            else {
                if (st != SLICE_I && st != SLICE_SI) prevMbSkipped = F.mb_skip_flag;
                if (mbaff && cur % 2 == 0) moreData = 1;
                else { PROF_T(t4); moreData = !cb.terminate(); PROF_ADD(12, t4); }
            }
            cur = next_mb(cur);
            if (cur < 0) break;
        } while (moreData);
        return 0;
    }
    int cabac_mb_field_flag_pre() { mbs[cur].slice = (uint16_t)F.slice_number; return cabac_mb_field_flag(); }
};

}  // namespace

int Front::decode_slice() {
    Dec d(*this);
    PROF_T(t0);
    const int r = d.run();
    PROF_ADD(6, t0);
    return r;
}

}  // namespace h264b2
