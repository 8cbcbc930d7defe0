// deblock.cuh — in-loop deblocking filter.
//
// Two kernels per batch:
//   k_bs       fully parallel: boundary strength for every (MB, edge, sample line), packed 4 bit each.
//              Mirrors DB:994-1310 incl. the reference's deviations (Q5 missing parentheses, Q6 raw
//              ref identity, Q13 left-edge p0 selection, Q14 unavailable-neighbour edges).
//   k_deblock  macroblock wavefront, in place: one warp owns one MB row (pair row under MBAFF) and
//              keeps the normative per-MB order (vertical edges left to right, then horizontal edges top
//              to bottom; MBs in address order, DB:76-635).  Lane = one sample line: lanes 0-15 luma,
//              16-23 Cb, 24-31 Cr (the three planes are independent).
#pragma once
#include "common.cuh"
#include "intra.cuh"   // flag helpers

__device__ const uint8_t g_alpha_tab[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
__device__ const uint8_t g_beta_tab[52]  = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18};
__device__ const uint8_t g_tc0_tab[3][52] = {
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,4,5,6,6,7,8,9,10,11,13},
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,5,5,6,7,8,8,10,11,12,13,15,17},
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,4,5,6,6,7,8,9,10,11,13,14,16,18,20,23,25}};

// thresholds of one (plane, edge kind): alpha, beta and tC0 for bS = 1, 2, 3 (DB:1314-1409) in one word, so the filter's
// dependent chain holds no table load
struct DbThr { int alpha, beta; uint32_t tc0; };      // tc0: 3 x 5 bits
__device__ __forceinline__ uint32_t db_thr_pack(int ia, int ib) {
    return (uint32_t)g_alpha_tab[ia] | ((uint32_t)g_beta_tab[ib] << 8) | ((uint32_t)g_tc0_tab[0][ia] << 13) | ((uint32_t)g_tc0_tab[1][ia] << 18) | ((uint32_t)g_tc0_tab[2][ia] << 23);
}
__device__ __forceinline__ DbThr db_thr_unpack(uint32_t w) {      // packed by k_bs
    DbThr t; t.alpha = w & 0xff; t.beta = (w >> 8) & 0x1f; t.tc0 = w >> 13; return t;
}

struct DbCtx { int A, B, left, top, leftflag, dbltop, fieldInFrame, internal, t8; };

__device__ inline DbCtx db_ctx(const PicDev &P, int a, const H264B2MbInfo &I) {   // DB:104-154
    DbCtx c;
    const int mbaff = P.mbaff, wmb = P.wmb;
    const int field = (I.flags & H264B2_MBF_FIELD) != 0, idc = I.deblock_idc;
    int xW, yW;
    c.A = nbr_loc(P, a, -1, 0, 0, xW, yW);
    c.B = nbr_loc(P, a, 0, -1, 0, xW, yW);
    c.t8 = (I.flags & H264B2_MBF_T8x8) != 0;
    c.fieldInFrame = mbaff && field;
    c.internal = idc != 1;
    c.left = !((!mbaff && a % wmb == 0) || (mbaff && (a >> 1) % wmb == 0) || idc == 1 || (idc == 2 && c.A < 0));
    c.top = !((!mbaff && a < wmb) || (mbaff && (a >> 1) < wmb && field) || (mbaff && (a >> 1) < wmb && !field && !(a & 1)) || idc == 1 || (idc == 2 && c.B < 0));
    c.leftflag = mbaff && a >= 2 && !field && (P.info[a - 2].flags & H264B2_MBF_FIELD);
    c.dbltop = mbaff && !(a & 1) && a >= 2 * wmb && !field && (P.info[a - 2 * wmb + 1].flags & H264B2_MBF_FIELD);
    return c;
}

__device__ __forceinline__ int is_intra_mode(const H264B2MbInfo &I) { return I.mb_class >= H264B2_MB_I4x4 && I.mb_class <= H264B2_MB_I16x16; }   // Q11

template <bool MBAFF>
__device__ __forceinline__ int derive_bs_core(const PicDev &P, const H264B2MbInfo &Ip, const H264B2MbInfo &Iq, int ap, int aq, int xp, int yp, int xq, int yq, int vertical) {   // DB:994
    const int mbaff = MBAFF;
    const int fp = (Ip.flags & H264B2_MBF_FIELD) != 0, fq = (Iq.flags & H264B2_MBF_FIELD) != 0;
    const int mixed = mbaff && ap != aq && fp != fq;
    const int intra = is_intra_mode(Ip) || is_intra_mode(Iq);
    const int spsi = ((Ip.flags | Iq.flags) & H264B2_MBF_SPSI) != 0;
    if (ap != aq) {
        if ((!fp && !fq && (intra || spsi)) || (mbaff && vertical && (intra || spsi))) return 4;
    }
    if ((!mixed && (intra || spsi)) || (mixed && !vertical && (intra || spsi))) return 3;
    const int bp = 8 * (yp / 8) + 4 * (xp / 8) + 2 * ((yp % 8) / 4) + ((xp % 8) / 4);
    const int bq = 8 * (yq / 8) + 4 * (xq / 8) + 2 * ((yq % 8) / 4) + ((xq % 8) / 4);
    if (((Ip.nnz_mask >> bp) & 1) || ((Iq.nnz_mask >> bq) & 1)) return 2;
    if (mixed) return 1;
    if (!P.motion) return 0;
    const H264B2MbMotion &Mp = P.motion[ap], &Mq = P.motion[aq];
    const int qp_ = (yp / 8) * 2 + xp / 8, qq_ = (yq / 8) * 2 + xq / 8, rp = (yp / 4) * 4 + xp / 4, rq = (yq / 4) * 4 + xq / 4;
    const int r0p = Mp.ref_ident[0][qp_], r1p = Mp.ref_ident[1][qp_], r0q = Mq.ref_ident[0][qq_], r1q = Mq.ref_ident[1][qq_];
    const int f0p = Mp.ref_surf[0][qp_] >= 0, f1p = Mp.ref_surf[1][qp_] >= 0, f0q = Mq.ref_surf[0][qq_] >= 0, f1q = Mq.ref_surf[1][qq_] >= 0;
    if (!(((r0p == r0q && r1p == r1q) || (r0p == r1q && r1p == r0q)) && (f0p + f1p) == (f0q + f1q))) return 1;
    const int lim = (mbaff && fq) ? 2 : 4;                                      // DB:1157
    const int m0px = Mp.mv[0][rp][0], m0py = Mp.mv[0][rp][1], m1px = Mp.mv[1][rp][0], m1py = Mp.mv[1][rp][1];
    const int m0qx = Mq.mv[0][rq][0], m0qy = Mq.mv[0][rq][1], m1qx = Mq.mv[1][rq][0], m1qy = Mq.mv[1][rq][1];
#define FAR(ax, ay, bx, by) (abs((ax) - (bx)) >= 4 || abs((ay) - (by)) >= lim)
    if (f0p && !f1p && f0q && !f1q && FAR(m0px, m0py, m0qx, m0qy)) return 1;
    if (f0p && !f1p && !f0q && f1q && FAR(m0px, m0py, m1qx, m1qy)) return 1;
    if (!f0p && f1p && f0q && !f1q && FAR(m1px, m1py, m0qx, m0qy)) return 1;
    if (!f0p && f1p && !f0q && f1q && FAR(m1px, m1py, m1qx, m1qy)) return 1;
    if (f0p && f1p && r0p != r1p && f0q && f1q && ((r0q == r0p && r1q == r1p) || (r0q == r1p && r1q == r0p))) {
        if (r0q == r0p && (FAR(m0px, m0py, m0qx, m0qy) || FAR(m1px, m1py, m1qx, m1qy))) return 1;
        else if (r0q == r1p && (FAR(m1px, m1py, m0qx, m0qy) || FAR(m0px, m0py, m1qx, m1qy))) return 1;
    }
    if (f0p && f1p && r0p == r1p && f0q && f1q && r0q == r1q && r0q == r0p) {
        // DB:1289-1298: A || (B && C) || D (missing parentheses in the reference, Q5)
        const int A = FAR(m0px, m0py, m0qx, m0qy), B = FAR(m1px, m1py, m1qx, m1qy), C = FAR(m0px, m0py, m1qx, m1qy), D = FAR(m1px, m1py, m0qx, m0qy);
        if (A || (B && C) || D) return 1;
    }
#undef FAR
    return 0;
}
__device__ inline int derive_bs(const PicDev &P, int ap, int aq, int xp, int yp, int xq, int yq, int vertical) {
    const H264B2MbInfo Ip = P.info[ap], Iq = P.info[aq];
    return P.mbaff ? derive_bs_core<true>(P, Ip, Iq, ap, aq, xp, yp, xq, yq, vertical) : derive_bs_core<false>(P, Ip, Iq, ap, aq, xp, yp, xq, yq, vertical);
}

// p0-side macroblock of sample line k of one edge (DB:742-805) and the bS coordinates (component units)
__device__ __forceinline__ int edge_p_mb(int a, int nbr, int vertical, int leftflag, int e, int k) {
    if (vertical) { if (e == 0 && nbr >= 0) return leftflag ? nbr + (k % 2) : nbr; return a; }
    const int yp = (e - 1) - (e % 2);
    if (yp < 0 && nbr >= 0) return nbr;
    return a;
}
__device__ inline int edge_bs(const PicDev &P, int a, int comp, int nbr, int vertical, int leftflag, int e, int k) {   // DB:639 + DB:859
    const int size = comp ? 8 : 16, s = comp ? 2 : 1;
    const int xE = vertical ? e : k, yE = vertical ? k : e;
    const int ap = edge_p_mb(a, nbr, vertical, leftflag, e, k);
    int xp, yp, xq, yq;
    if (vertical) { xp = xE - 1; if (xp < 0) xp += size; yp = yE; xq = xE; yq = yE; }
    else { xp = xE; yp = (yE - 1) - (yE % 2); if (yp < 0) yp += size; xq = xE; yq = yE - (yE % 2); }
    return derive_bs(P, ap, a, xp * s, yp * s, xq * s, yq * s, vertical);
}

// Progressive pictures: one warp per macroblock, lane = (direction, edge, 4-sample segment).  2-D grid (no division), the
// three info records a macroblock's edges can touch are loaded once per warp, and the strength derivation is the
// frame-only instance of derive_bs_core.  Record layout: see k_bs below.
__global__ void __launch_bounds__(256) k_bs_prog(const PicDev *pics) {
    const PicDev &P = pics[blockIdx.z];
    const int lane = threadIdx.x & 31;
    const int mbx = blockIdx.x * 8 + (threadIdx.x >> 5), mby = blockIdx.y, wmb = P.wmb;
    if (mbx >= wmb) return;
    const int a = mby * wmb + mbx, nmb = wmb * P.hmb;
    if (!P.deblock_enable || a >= P.deblock_stop || P.generic) return;
    const H264B2MbInfo I = P.info[a];
    H264B2MbInfo IA = I, IB = I;
    int A = -1, B = -1;                                                         // DB:13-69 / PB:2878: same slice, address <= current
    if (mbx > 0) { IA = P.info[a - 1]; if (IA.slice_number == I.slice_number) A = a - 1; }
    if (mby > 0) { IB = P.info[a - wmb]; if (IB.slice_number == I.slice_number) B = a - wmb; }
    const int idc = I.deblock_idc, t8 = (I.flags & H264B2_MBF_T8x8) != 0;
    const int left = !(mbx == 0 || idc == 1 || (idc == 2 && A < 0)), top = !(mby == 0 || idc == 1 || (idc == 2 && B < 0)), internal = idc != 1;   // DB:104-154
    const int dir = lane >> 4, edge = (lane >> 2) & 3, seg = lane & 3;
    const int on = edge == 0 ? (dir ? top : left) : (internal && (!t8 || edge == 2));
    int bS = 0;
    if (on) {
        // DB:639-840 for frame macroblocks: q = this MB at the edge, p = the sample before it (Q14: the MB itself, far side, when
        // the neighbour is not available)
        const int nbr = dir ? B : A;
        const bool outer = edge == 0 && nbr >= 0;
        const int ap = outer ? nbr : a;
        const int xq = dir ? 4 * seg : 4 * edge, yq = dir ? 4 * edge : 4 * seg;
        const int xp = dir ? xq : (edge ? xq - 1 : 15), yp = dir ? (edge ? yq - 1 : 15) : yq;
        bS = derive_bs_core<false>(P, outer ? (dir ? IB : IA) : I, I, ap, a, xp, yp, xq, yq, !dir);
    }
    const int idx = lane & 15;
    const uint32_t val = (uint32_t)bS << (4 * (idx & 7));
    const int word = dir * 2 + (idx >> 3);
    const uint32_t r0 = __reduce_or_sync(0xffffffffu, word == 0 ? val : 0u), r1 = __reduce_or_sync(0xffffffffu, word == 1 ? val : 0u);
    const uint32_t r2 = __reduce_or_sync(0xffffffffu, word == 2 ? val : 0u), r3 = __reduce_or_sync(0xffffffffu, word == 3 ? val : 0u);
    uint32_t *rec = P.bs + (size_t)a * 16;
    if (lane == 0) { *(uint4 *)rec = make_uint4(r0, r1, r2, r3); P.bs[(size_t)nmb * 64 + a] = (r0 | r1 | r2 | r3) != 0; }
    if (lane < 9 && (r0 | r1 | r2 | r3)) {
        const int cc = lane / 3, t = lane % 3;
        int qq = I.mb_class == H264B2_MB_IPCM ? 0 : I.qpy, qp = qq;
        if (t == 0 && A >= 0) qp = IA.mb_class == H264B2_MB_IPCM ? 0 : IA.qpy;
        if (t == 1 && B >= 0) qp = IB.mb_class == H264B2_MB_IPCM ? 0 : IB.qpy;
        if (cc) { qq = chroma_qp(P, qq, cc - 1); qp = chroma_qp(P, qp, cc - 1); }
        const int qpav = (qp + qq + 1) >> 1;
        rec[4 + lane] = db_thr_pack(clip3i(0, 51, qpav + I.filter_offset_a), clip3i(0, 51, qpav + I.filter_offset_b));
    }
}

// packed layout per MB: word[lane] = vertical edges (4 bit per edge index), word[32+lane] = horizontal
// edge slots (0: top edge / first field pass, 1: second field pass of a frame MB under a field pair,
// 2..4: internal edges 4, 8, 12).  word[64*n_mbs + a] = 1 when any strength of MB a is non-zero.
__global__ void __launch_bounds__(256) k_bs(const PicDev *pics) {
    const PicDev &P = pics[blockIdx.y];
    const int a = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int nmb = P.wmb * P.hmb;
    if (a >= nmb || !P.deblock_enable || a >= P.deblock_stop) return;
    const H264B2MbInfo I = P.info[a];
    const DbCtx c = db_ctx(P, a, I);
    if (!P.generic) {
        // ---- compact record for progressive pictures (consumed by deblock_row_fast): without MBAFF the strength
        // is constant over each 4-sample segment and chroma line k reuses luma line 2k (DB:871-880), so 16 vertical
        // + 16 horizontal values describe the MB.  16 words per MB at bs[a*16]:
        //   [0..1] vertical nibbles (index edge*4 + segment), [2..3] horizontal nibbles,
        //   [4 + comp*3 + t] thresholds (db_thr_pack: alpha, beta, tC0 of bS 1..3) for t = 0 left edge, 1 top edge, 2 internal edges
        const int dir = lane >> 4, edge = (lane >> 2) & 3, seg = lane & 3;
        const int on = edge == 0 ? (dir ? c.top : c.left) : (c.internal && (!c.t8 || edge == 2));
        const int bS = on ? edge_bs(P, a, 0, dir ? c.B : c.A, !dir, 0, 4 * edge, 4 * seg) : 0;
        const int idx = lane & 15;
        const uint32_t val = (uint32_t)bS << (4 * (idx & 7));
        const int word = dir * 2 + (idx >> 3);
        uint32_t r0 = __reduce_or_sync(0xffffffffu, word == 0 ? val : 0u), r1 = __reduce_or_sync(0xffffffffu, word == 1 ? val : 0u);
        uint32_t r2 = __reduce_or_sync(0xffffffffu, word == 2 ? val : 0u), r3 = __reduce_or_sync(0xffffffffu, word == 3 ? val : 0u);
        uint32_t *rec = P.bs + (size_t)a * 16;
        if (lane == 0) { *(uint4 *)rec = make_uint4(r0, r1, r2, r3); P.bs[(size_t)nmb * 64 + a] = (r0 | r1 | r2 | r3) != 0; }
        if (lane < 9) {
            const int cc = lane / 3, t = lane % 3;
            int qq = I.mb_class == H264B2_MB_IPCM ? 0 : I.qpy, qp = qq;
            const int n = t == 0 ? c.A : t == 1 ? c.B : -1;
            if (n >= 0) { const H264B2MbInfo In = P.info[n]; qp = In.mb_class == H264B2_MB_IPCM ? 0 : In.qpy; }
            if (cc) { qq = chroma_qp(P, qq, cc - 1); qp = chroma_qp(P, qp, cc - 1); }
            const int qpav = (qp + qq + 1) >> 1;
            const int ia = clip3i(0, 51, qpav + I.filter_offset_a), ib = clip3i(0, 51, qpav + I.filter_offset_b);
            rec[4 + lane] = db_thr_pack(ia, ib);
        }
        return;
    }
    const int comp = lane < 16 ? 0 : lane < 24 ? 1 : 2;
    const int k = comp ? (lane & 7) : lane;
    const int ne = comp ? 2 : 4;
    uint32_t v = 0, h = 0;
    for (int i = 0; i < ne; i++) {
        const int e = 4 * i;
        const int on = i == 0 ? c.left : (c.internal && (comp || !c.t8 || i == 2));
        if (on) v |= (uint32_t)edge_bs(P, a, comp, c.A, 1, i == 0 ? c.leftflag : 0, e, k) << (4 * i);
    }
    if (c.top) {
        if (c.dbltop) {
            h |= (uint32_t)edge_bs(P, a, comp, comp ? c.B : c.B - 1, 0, 0, 0, k);       // chroma uses mbAddrB in both passes (DB:474-497)
            h |= (uint32_t)edge_bs(P, a, comp, c.B, 0, 0, 1, k) << 4;
        } else h |= (uint32_t)edge_bs(P, a, comp, c.B, 0, 0, 0, k);
    }
    if (c.internal) for (int i = 1; i < ne; i++) {
        if (comp || !c.t8 || i == 2) h |= (uint32_t)edge_bs(P, a, comp, c.B, 0, 0, 4 * i, k) << (4 * (i + 1));
    }
    P.bs[(size_t)a * 64 + lane] = v;
    P.bs[(size_t)a * 64 + 32 + lane] = h;
    const unsigned any = __ballot_sync(0xffffffffu, (v | h) != 0);
    if (lane == 0) P.bs[(size_t)nmb * 64 + a] = any != 0;
}

// filter one sample line across one edge (DB:844 tail, DB:1314-1522).  q0 points at q0; p_i = q0[-(i+1)*step].
__device__ inline void filter_line(const PicDev &P, int comp, int bS, int ap, int aq, uint8_t *q0, int step) {
    int p[4], q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { q[i] = __ldcg(q0 + i * step); p[i] = __ldcg(q0 - (i + 1) * step); }
    const H264B2MbInfo Ip = P.info[ap], Iq = P.info[aq];
    int qpp = Ip.mb_class == H264B2_MB_IPCM ? 0 : Ip.qpy, qpq = Iq.mb_class == H264B2_MB_IPCM ? 0 : Iq.qpy;
    if (comp) { qpp = chroma_qp(P, qpp, comp - 1); qpq = chroma_qp(P, qpq, comp - 1); }
    const int qpav = (qpp + qpq + 1) >> 1;
    const int ia = clip3i(0, 51, qpav + Iq.filter_offset_a), ib = clip3i(0, 51, qpav + Iq.filter_offset_b);
    const int alpha = g_alpha_tab[ia], beta = g_beta_tab[ib];
    if (!(abs(p[0] - q[0]) < alpha && abs(p[1] - p[0]) < beta && abs(q[1] - q[0]) < beta)) return;
    const int ap_ = abs(p[2] - p[0]), aq_ = abs(q[2] - q[0]);
    int np0 = p[0], np1 = p[1], np2 = p[2], nq0 = q[0], nq1 = q[1], nq2 = q[2];
    if (bS < 4) {
        const int tc0 = g_tc0_tab[bS - 1][ia];
        const int tc = comp ? tc0 + 1 : tc0 + (ap_ < beta) + (aq_ < beta);
        const int delta = clip3i(-tc, tc, (((q[0] - p[0]) << 2) + (p[1] - q[1]) + 4) >> 3);
        np0 = clip255(p[0] + delta); nq0 = clip255(q[0] - delta);
        if (!comp && ap_ < beta) np1 = p[1] + clip3i(-tc0, tc0, (p[2] + ((p[0] + q[0] + 1) >> 1) - (p[1] << 1)) >> 1);
        if (!comp && aq_ < beta) nq1 = q[1] + clip3i(-tc0, tc0, (q[2] + ((p[0] + q[0] + 1) >> 1) - (q[1] << 1)) >> 1);
    } else {
        const int small = abs(p[0] - q[0]) < ((alpha >> 2) + 2);
        if (!comp && ap_ < beta && small) { np0 = (p[2] + 2*p[1] + 2*p[0] + 2*q[0] + q[1] + 4) >> 3; np1 = (p[2] + p[1] + p[0] + q[0] + 2) >> 2; np2 = (2*p[3] + 3*p[2] + p[1] + p[0] + q[0] + 4) >> 3; }
        else np0 = (2*p[1] + p[0] + q[1] + 2) >> 2;
        if (!comp && aq_ < beta && small) { nq0 = (p[1] + 2*p[0] + 2*q[0] + 2*q[1] + q[2] + 4) >> 3; nq1 = (p[0] + q[0] + q[1] + q[2] + 2) >> 2; nq2 = (2*q[3] + 3*q[2] + q[1] + q[0] + p[0] + 4) >> 3; }
        else nq0 = (2*q[1] + q[0] + p[1] + 2) >> 2;
    }
    q0[0] = (uint8_t)nq0; q0[step] = (uint8_t)nq1; q0[2 * step] = (uint8_t)nq2;
    q0[-step] = (uint8_t)np0; q0[-2 * step] = (uint8_t)np1; q0[-3 * step] = (uint8_t)np2;
}

__device__ inline void deblock_mb(const PicDev &P, int a, int lane) {
    const int nmb = P.wmb * P.hmb;
    const uint32_t v = P.bs[(size_t)a * 64 + lane], h = P.bs[(size_t)a * 64 + 32 + lane];
    const H264B2MbInfo I = P.info[a];
    const DbCtx c = db_ctx(P, a, I);
    const int comp = lane < 16 ? 0 : lane < 24 ? 1 : 2;
    const int k = comp ? (lane & 7) : lane;
    const int W = P.wmb * 16, H = P.hmb * 16;
    const int stride = comp ? W >> 1 : W;
    uint8_t *pl = P.dst + (comp ? (size_t)W * H + (comp == 2 ? (size_t)(W >> 1) * (H >> 1) : 0) : 0);
    int xI, yI;
    mb_origin(P, a, c.fieldInFrame, xI, yI);
    const int xP = comp ? xI / 2 : xI, yP = comp ? (yI + 1) / 2 : yI;
    (void)nmb;
    // vertical edges: lane = row k
    if (v) {
        const int dy = 1 + c.fieldInFrame;
        uint8_t *rowp = pl + (size_t)(yP + dy * k) * stride + xP;
        const int ne = comp ? 2 : 4;
        for (int i = 0; i < ne; i++) {
            const int bS = (v >> (4 * i)) & 15;
            if (bS) filter_line(P, comp, bS, edge_p_mb(a, c.A, 1, i == 0 ? c.leftflag : 0, 4 * i, k), a, rowp + 4 * i, 1);
        }
    }
    __syncwarp();
    // horizontal edges: lane = column k
    if (h) {
        for (int slot = 0; slot < 5; slot++) {
            const int bS = (h >> (4 * slot)) & 15;
            if (!bS) continue;
            int e, fieldmode, nbr;
            if (slot == 0) { e = 0; fieldmode = c.dbltop ? 1 : c.fieldInFrame; nbr = c.dbltop ? (comp ? c.B : c.B - 1) : c.B; }
            else if (slot == 1) { e = 1; fieldmode = 1; nbr = c.B; }
            else { e = 4 * (slot - 1); fieldmode = c.fieldInFrame; nbr = c.B; }
            const int dy = 1 + fieldmode;
            uint8_t *q0 = pl + (size_t)(yP + dy * e - (e % 2)) * stride + xP + k;
            filter_line(P, comp, bS, edge_p_mb(a, nbr, 0, 0, e, k), a, q0, dy * stride);
        }
    }
    __syncwarp();
}

#include "deblock_fast.cuh"

// GENERIC = false: every picture of the launch is progressive (PicDev::generic == 0) -> only the staged fast path is
// compiled in (smaller code, fewer registers); GENERIC = true: MBAFF pictures, literal per-sample-line walk.
template <bool GENERIC>
__global__ void __launch_bounds__(WF_THREADS, 1024 / WF_THREADS) k_deblock(const PicDev *pics, int npics, int bands, int *ticket) {
    __shared__ DbTile tiles[WF_ROWS];
    __shared__ int s_prog[WF_ROWS];
    __shared__ uint64_t s_bar[WF_ROWS];
    __shared__ int s_ticket;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    if (threadIdx.x < WF_ROWS) { s_prog[threadIdx.x] = 0; mbar_init(&s_bar[threadIdx.x], 1); }
    __syncthreads();
    const int t = s_ticket;
    if (t >= npics * bands) return;
    const PicDev &P = pics[t % npics];
    if (!P.deblock_enable) return;
    const int row = (t / npics) * WF_ROWS + warp;
    const int per = P.mbaff ? 2 : 1;
    const int rows = P.hmb / per, wmb = P.wmb, nmb = P.wmb * P.hmb;
    if (row >= rows) return;
    RowSync rs = rs_init(s_prog, s_bar, warp, row, rows, P.progress + P.hmb, wmb);     // progress[1][row]
    if (!GENERIC) { deblock_row_fast(P, row, lane, tiles[warp], rs); return; }
    const uint32_t *anyflag = P.bs + (size_t)nmb * 64;
    for (int xb = 0; xb < wmb; xb += 32) {
        const int xl = xb + lane;
        int work = 0;
        if (xl < wmb) for (int s = 0; s < per; s++) { const int a = (row * wmb + xl) * per + s; if (a < P.deblock_stop) work |= anyflag[a] != 0; }
        unsigned mask = __ballot_sync(0xffffffffu, work);
        while (mask) {
            const int x = xb + __ffs(mask) - 1;
            mask &= mask - 1;
            rs_wait(rs, min(x + 2, wmb), x, lane);
            for (int s = 0; s < per; s++) {
                const int a = (row * wmb + x) * per + s;
                if (a < P.deblock_stop && anyflag[a]) {
                    // frame MB whose left and above pairs are frame pairs: staged filter; else the literal DB:639 walk
                    const H264B2MbInfo I = P.info[a];
                    const DbCtx c = db_ctx(P, a, I);
                    const int w = P.wmb, pr = a >> 1;
                    bool tile = P.mbaff && !(I.flags & H264B2_MBF_FIELD) && !c.leftflag && !c.dbltop;
                    if (tile && pr % w > 0) tile = !(P.info[a - 2].flags & H264B2_MBF_FIELD);
                    if (tile && pr >= w) tile = !(P.info[2 * (pr - w)].flags & H264B2_MBF_FIELD);
                    if (tile) deblock_mb_tile(P, a, c.A, c.B, lane, tiles[warp]); else deblock_mb(P, a, lane);
                }
            }
            rs_publish(rs, x + 1, lane);
        }
    }
    rs_publish(rs, wmb, lane);
}
