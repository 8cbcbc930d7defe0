// deblock_fast.cuh — the deblocking wavefront for progressive (non-MBAFF) pictures, staged in shared memory.
//
// Same normative order as the generic path in deblock.cuh (per MB: vertical edges 0,4,8,12 then horizontal
// edges 0,4,8,12; MBs left to right; rows as a 2:1 wavefront), but one warp keeps its current macroblock
// in a shared-memory tile instead of going to L2 for every sample line:
//   * the MB's own 16x16 / 8x8 / 8x8 samples are PREFETCHED into registers one MB ahead (nobody modifies
//     them before this warp does), so their latency hides behind the previous MB's filtering;
//   * the right-hand 4 (luma) / 4 (chroma) columns stay in the tile as the next MB's left neighbour;
//   * only the 4 luma / 2 chroma rows of the MB above come from L2 after the row-above flag is acquired;
//   * a sample line is two 32-bit words (p3..p0 | q0..q3): the same routine filters rows (words as loaded)
//     and columns (words gathered bytewise from the tile);
//   * progress is published with st.release after the MB's rows are written back.
#pragma once
#include "common.cuh"
#include "wavefront.cuh"

struct DbTile {
    uint32_t L[20][5];        // luma rows -4..15 (index r+4); word 0 = columns -4..-1, words 1..4 = columns 0..15
    uint32_t C[2][12][3];     // Cb, Cr rows -4..7 (index r+4; rows -4,-3 unused); word 0 = columns -4..-1, words 1..2 = columns 0..7
    uint32_t work[8];         // bit x: macroblock x of this row has at least one non-zero boundary strength
};

__device__ __forceinline__ void st_release_flag(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// one sample line: Pw = p3 p2 p1 p0 (byte 3 = p0), Qw = q0 q1 q2 q3 (byte 0 = q0).  DB:1373 / DB:1481.
#ifndef DB_FILTER_NOINLINE
#define DB_FILTER_NOINLINE 0
#endif
__device__ __forceinline__ void filter_words_body(uint32_t &Pw, uint32_t &Qw, int bS, const DbThr &t, int chroma) {
    const int p0 = Pw >> 24, p1 = (Pw >> 16) & 0xff, p2 = (Pw >> 8) & 0xff, p3 = Pw & 0xff;
    const int q0 = Qw & 0xff, q1 = (Qw >> 8) & 0xff, q2 = (Qw >> 16) & 0xff, q3 = Qw >> 24;
    if (!(abs(p0 - q0) < t.alpha && abs(p1 - p0) < t.beta && abs(q1 - q0) < t.beta)) return;
    int np0 = p0, np1 = p1, np2 = p2, nq0 = q0, nq1 = q1, nq2 = q2;
    const int ap = abs(p2 - p0), aq = abs(q2 - q0);
    if (bS < 4) {
        const int tc0 = (int)((t.tc0 >> (5 * (bS - 1))) & 31u);
        const int tc = chroma ? tc0 + 1 : tc0 + (ap < t.beta) + (aq < t.beta);
        const int delta = clip3i(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        np0 = clip255(p0 + delta); nq0 = clip255(q0 - delta);
        if (!chroma && ap < t.beta) np1 = p1 + clip3i(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1);
        if (!chroma && aq < t.beta) nq1 = q1 + clip3i(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1);
    } else {
        const int small = abs(p0 - q0) < ((t.alpha >> 2) + 2);
        if (!chroma && ap < t.beta && small) { np0 = (p2 + 2*p1 + 2*p0 + 2*q0 + q1 + 4) >> 3; np1 = (p2 + p1 + p0 + q0 + 2) >> 2; np2 = (2*p3 + 3*p2 + p1 + p0 + q0 + 4) >> 3; }
        else np0 = (2*p1 + p0 + q1 + 2) >> 2;
        if (!chroma && aq < t.beta && small) { nq0 = (p1 + 2*p0 + 2*q0 + 2*q1 + q2 + 4) >> 3; nq1 = (p0 + q0 + q1 + q2 + 2) >> 2; nq2 = (2*q3 + 3*q2 + q1 + q0 + p0 + 4) >> 3; }
        else nq0 = (2*q1 + q0 + p1 + 2) >> 2;
    }
    Pw = (uint32_t)p3 | ((uint32_t)np2 << 8) | ((uint32_t)np1 << 16) | ((uint32_t)np0 << 24);
    Qw = (uint32_t)nq0 | ((uint32_t)nq1 << 8) | ((uint32_t)nq2 << 16) | ((uint32_t)q3 << 24);
}

// ONE copy of the filter in the kernel (DB_FILTER_NOINLINE=1): k_deblock's eight inlined copies were 44 KB of SASS and ncu showed
// instruction-fetch stalls; the call passes everything in registers (packed words in, packed words out).
__device__ __noinline__ uint2 filter_words_call(uint32_t Pw, uint32_t Qw, int bS, uint32_t thr, int chroma) {
    const DbThr t = db_thr_unpack(thr);
    filter_words_body(Pw, Qw, bS, t, chroma);
    return make_uint2(Pw, Qw);
}
__device__ __forceinline__ uint32_t db_thr_repack(const DbThr &t) { return (uint32_t)t.alpha | ((uint32_t)t.beta << 8) | (t.tc0 << 13); }
__device__ __forceinline__ void filter_words(uint32_t &Pw, uint32_t &Qw, int bS, const DbThr &t, int chroma) {
#if DB_FILTER_NOINLINE
    const uint2 r = filter_words_call(Pw, Qw, bS, db_thr_repack(t), chroma);
    Pw = r.x; Qw = r.y;
#else
    filter_words_body(Pw, Qw, bS, t, chroma);
#endif
}

// vertical phase on the staged tile: lane = sample row; step i filters the edge between tile words i and i+1
__device__ __forceinline__ void db_vertical(DbTile &T, int comp, int k, uint32_t v, const DbThr &tA, const DbThr &tI) {
    uint32_t *rowp = comp ? &T.C[comp - 1][4 + k][0] : &T.L[4 + k][0];
    if (v) {
        uint32_t w0 = rowp[0], w1 = rowp[1], w2 = rowp[2], w3 = 0, w4 = 0;
        if (!comp) { w3 = rowp[3]; w4 = rowp[4]; }
        if (v & 0xF) filter_words(w0, w1, v & 15, tA, comp);
        if (v & 0xF0) filter_words(w1, w2, (v >> 4) & 15, tI, comp);
        if (v & 0xF00) filter_words(w2, w3, (v >> 8) & 15, tI, comp);
        if (v & 0xF000) filter_words(w3, w4, (v >> 12) & 15, tI, comp);
        rowp[0] = w0; rowp[1] = w1; rowp[2] = w2;
        if (!comp) { rowp[3] = w3; rowp[4] = w4; }
    }
}
// horizontal phase: lane = sample column; the column is gathered into words of 4 rows, then the same 4 steps
__device__ __forceinline__ void db_horizontal(DbTile &T, int comp, int k, uint32_t h, const DbThr &tB, const DbThr &tI) {
    if (h) {
        uint8_t *col = comp ? (uint8_t *)&T.C[comp - 1][0][0] + 4 + k : (uint8_t *)&T.L[0][0] + 4 + k;
        const int sb = comp ? 12 : 20;
        uint32_t cw[5];
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if (j < 3 || !comp) cw[j] = (uint32_t)col[(4 * j) * sb] | ((uint32_t)col[(4 * j + 1) * sb] << 8) | ((uint32_t)col[(4 * j + 2) * sb] << 16) | ((uint32_t)col[(4 * j + 3) * sb] << 24);
            else cw[j] = 0;
        }
        if (h & 0xF) filter_words(cw[0], cw[1], h & 15, tB, comp);
        if (h & 0xF0) filter_words(cw[1], cw[2], (h >> 4) & 15, tI, comp);
        if (h & 0xF00) filter_words(cw[2], cw[3], (h >> 8) & 15, tI, comp);
        if (h & 0xF000) filter_words(cw[3], cw[4], (h >> 12) & 15, tI, comp);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if (j < 3 || !comp) { col[(4 * j) * sb] = (uint8_t)cw[j]; col[(4 * j + 1) * sb] = (uint8_t)(cw[j] >> 8); col[(4 * j + 2) * sb] = (uint8_t)(cw[j] >> 16); col[(4 * j + 3) * sb] = (uint8_t)(cw[j] >> 24); }
        }
    }
}

__device__ __forceinline__ int db_next_work(const uint32_t *work, int x, int wmb) {   // first work MB with index > x, or wmb
    for (int i = x + 1; i < wmb; ) {
        const uint32_t w = work[i >> 5] >> (i & 31);
        if (w) return i + __ffs(w) - 1;
        i = (i | 31) + 1;
    }
    return wmb;
}

struct DbPrefetch { uint32_t y0, y1, c, l; uint4 bs; uint32_t tA, tB, tI; };   // own luma words (lane, lane+32), own chroma word, left-column word, k_bs record

// lanes: luma word w = lane + 32 t -> row w/4, word w%4; chroma: lanes 0-15 Cb word (row = l/2, word l%2), 16-31 Cr
__device__ __forceinline__ DbPrefetch db_prefetch(const PicDev &P, int row, int x, int lane, int need_left) {
    const int W = P.wmb * 16, H = P.hmb * 16, Wc = W >> 1;
    const uint8_t *Y = P.dst + (size_t)(row * 16) * W + x * 16;
    const uint8_t *C = P.dst + (size_t)W * H + (lane >= 16 ? (size_t)Wc * (H >> 1) : 0) + (size_t)(row * 8) * Wc + x * 8;
    DbPrefetch f;
    f.y0 = __ldcg((const uint32_t *)(Y + (size_t)(lane >> 2) * W + (lane & 3) * 4));
    f.y1 = __ldcg((const uint32_t *)(Y + (size_t)(8 + (lane >> 2)) * W + (lane & 3) * 4));
    const int cl = lane & 15;
    f.c = __ldcg((const uint32_t *)(C + (size_t)(cl >> 1) * Wc + (cl & 1) * 4));
    {   const int comp = lane < 16 ? 0 : lane < 24 ? 1 : 2;
        const uint32_t *rec = P.bs + (size_t)(row * P.wmb + x) * 16;
        f.bs = *(const uint4 *)rec;
        f.tA = rec[4 + comp * 3]; f.tB = rec[5 + comp * 3]; f.tI = rec[6 + comp * 3]; }
    f.l = 0;
    if (need_left && x > 0) {
        if (lane < 16) f.l = __ldcg((const uint32_t *)(Y + (size_t)lane * W - 4));
        else if (lane < 24) f.l = __ldcg((const uint32_t *)(P.dst + (size_t)W * H + (size_t)(row * 8 + lane - 16) * Wc + x * 8 - 4));
        else f.l = __ldcg((const uint32_t *)(P.dst + (size_t)W * H + (size_t)Wc * (H >> 1) + (size_t)(row * 8 + lane - 24) * Wc + x * 8 - 4));
    }
    return f;
}

// Deblock one MB row of a progressive picture.  prog = this picture's deblock progress counters.
__device__ inline void deblock_row_fast(const PicDev &P, int row, int lane, DbTile &T, RowSync &rs) {
    const int wmb = P.wmb, nmb = P.wmb * P.hmb;
    const int W = wmb * 16, H = P.hmb * 16, Wc = W >> 1;
    const uint32_t *anyflag = P.bs + (size_t)nmb * 64;
    for (int g = 0; g < 8; g++) {
        const int xl = g * 32 + lane;
        int work = 0;
        if (xl < wmb) { const int a = row * wmb + xl; if (a < P.deblock_stop) work = anyflag[a] != 0; }
        const unsigned m = __ballot_sync(0xffffffffu, work);
        if (lane == 0) T.work[g] = m;
    }
    __syncwarp();
    const int comp = lane < 16 ? 0 : lane < 24 ? 1 : 2;
    const int k = comp ? (lane & 7) : lane;
    int x = db_next_work(T.work, -1, wmb);
    int carry_x = -2;                                    // tile word 0 holds the right columns of MB carry_x
    DbPrefetch pf;
    if (x < wmb) pf = db_prefetch(P, row, x, lane, 1);
    while (x < wmb) {
        // ---- stage the prefetched MB into the tile
        T.L[4 + (lane >> 2)][1 + (lane & 3)] = pf.y0;
        T.L[12 + (lane >> 2)][1 + (lane & 3)] = pf.y1;
        { const int cl = lane & 15; T.C[lane >> 4][4 + (cl >> 1)][1 + (cl & 1)] = pf.c; }
        if (carry_x != x - 1 && x > 0) { if (lane < 16) T.L[4 + lane][0] = pf.l; else T.C[(lane - 16) >> 3][4 + ((lane - 16) & 7)][0] = pf.l; }
        // ---- per-MB parameters (prefetched k_bs record): this lane's strengths, one nibble per filter step.
        //      Luma and chroma lanes run the SAME 4-step sequence (chroma has strengths only in steps 0 and 1,
        //      which are its edges 0 and 4 = luma edges 0 and 8), so the warp never serialises the two planes.
        uint32_t v, h;
        {
            const unsigned long long bv = (unsigned long long)pf.bs.x | ((unsigned long long)pf.bs.y << 32);
            const unsigned long long bh = (unsigned long long)pf.bs.z | ((unsigned long long)pf.bs.w << 32);
            const int seg = comp ? k >> 1 : k >> 2;      // chroma line k <-> luma line 2k
            const int e1 = comp ? 8 : 4;                 // nibble index of the step-1 edge: luma edge 4, chroma edge 4 (= luma edge 8)
            v = (uint32_t)((bv >> (4 * seg)) & 15) | ((uint32_t)((bv >> (4 * (e1 + seg))) & 15) << 4);
            h = (uint32_t)((bh >> (4 * seg)) & 15) | ((uint32_t)((bh >> (4 * (e1 + seg))) & 15) << 4);
            if (!comp) {
                v |= ((uint32_t)((bv >> (4 * (8 + seg))) & 15) << 8) | ((uint32_t)((bv >> (4 * (12 + seg))) & 15) << 12);
                h |= ((uint32_t)((bh >> (4 * (8 + seg))) & 15) << 8) | ((uint32_t)((bh >> (4 * (12 + seg))) & 15) << 12);
            }
        }
        const DbThr tA = db_thr_unpack(pf.tA), tB = db_thr_unpack(pf.tB), tI = db_thr_unpack(pf.tI);
        const int xn = db_next_work(T.work, x, wmb);
        if (xn < wmb) pf = db_prefetch(P, row, xn, lane, xn != x + 1);
        const int topf = __any_sync(0xffffffffu, (h & 15u) != 0);
        // ---- rows of the MB above: only the horizontal phase needs them.  If the row above is already far enough,
        //      fetch them now so that the L2 latency hides behind the vertical phase; otherwise wait after it.
        const uint8_t *toprow = nullptr;
        if (lane < 16) toprow = P.dst + (size_t)(row * 16 - 4 + (lane >> 2)) * W + x * 16 + (lane & 3) * 4;
        else if (lane < 24) { const int l = lane - 16; toprow = P.dst + (size_t)W * H + ((l >> 2) ? (size_t)Wc * (H >> 1) : 0) + (size_t)(row * 8 - 2 + ((l >> 1) & 1)) * Wc + x * 8 + (l & 1) * 4; }
        uint32_t topw = 0;
        const bool early = topf && rs_try(rs, min(x + 2, wmb), lane);
        if (early && toprow) topw = __ldcg((const uint32_t *)toprow);
        __syncwarp();
        db_vertical(T, comp, k, v, tA, tI);
        if (topf) {
            if (!early) { rs_wait(rs, min(x + 2, wmb), x, lane); if (toprow) topw = __ldcg((const uint32_t *)toprow); }
            if (lane < 16) T.L[lane >> 2][1 + (lane & 3)] = topw;
            else if (lane < 24) { const int l = lane - 16; T.C[l >> 2][2 + ((l >> 1) & 1)][1 + (l & 1)] = topw; }
        }
        __syncwarp();
        db_horizontal(T, comp, k, h, tB, tI);
        __syncwarp();
        // ---- write back: columns -4..11 of the MB's rows (the left MB's last columns are final now), the
        //      rows above if they were filtered, and at the end of the row also columns 12..15
        {
            uint8_t *Y = P.dst + (size_t)(row * 16) * W + x * 16;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int w = lane + 32 * t, r = w >> 2, j = w & 3;
                if (j > 0 || x > 0) *(uint32_t *)(Y + (size_t)r * W + j * 4 - 4) = T.L[4 + r][j];
            }
            const int c = lane >> 4, cl = lane & 15, r = cl >> 1, j = cl & 1;
            uint8_t *Cp = P.dst + (size_t)W * H + (c ? (size_t)Wc * (H >> 1) : 0) + (size_t)(row * 8) * Wc + x * 8;
            if (j > 0 || x > 0) *(uint32_t *)(Cp + (size_t)r * Wc + j * 4 - 4) = T.C[c][4 + r][j];
            if (xn != x + 1 || xn >= wmb) {      // nobody will carry our right-hand columns: write them now
                if (lane < 16) *(uint32_t *)(Y + (size_t)lane * W + 12) = T.L[4 + lane][4];
                else { const int l = lane - 16, cc = l >> 3, rr = l & 7;
                    *(uint32_t *)(P.dst + (size_t)W * H + (cc ? (size_t)Wc * (H >> 1) : 0) + (size_t)(row * 8 + rr) * Wc + x * 8 + 4) = T.C[cc][4 + rr][2]; }
            }
            if (topf) {
                if (lane < 12) { const int rr = 1 + lane / 4, jj = lane & 3; *(uint32_t *)(Y + (size_t)(rr - 4) * W + jj * 4) = T.L[rr][1 + jj]; }
                else if (lane >= 16 && lane < 20) { const int l = lane - 16, cc = l >> 1, jj = l & 1;
                    *(uint32_t *)(P.dst + (size_t)W * H + (cc ? (size_t)Wc * (H >> 1) : 0) + (size_t)(row * 8 - 1) * Wc + x * 8 + jj * 4) = T.C[cc][3][1 + jj]; }
            }
        }
        // ---- carry the right-hand columns to the next MB
        if (lane < 16) T.L[4 + lane][0] = T.L[4 + lane][4];
        else { const int l = lane - 16; T.C[l >> 3][4 + (l & 7)][0] = T.C[l >> 3][4 + (l & 7)][2]; }
        carry_x = x;
        rs_publish(rs, x + 1, lane);
        x = xn;
    }
    rs_publish(rs, wmb, lane);
}

// One frame macroblock of an MBAFF picture whose left and above neighbour pairs are frame pairs too (no mixed edges,
// no field stepping): same staged filter, but every sample comes from / goes to global memory for this MB alone (the
// pair-interleaved address order rules out the row-wise carry).  v/h: per-line strengths written by k_bs for MBAFF
// pictures; A/B: left / top neighbour MB for the thresholds (-1 = unavailable -> the MB itself, Q14).
__device__ inline void deblock_mb_tile(const PicDev &P, int a, int A, int B, int lane, DbTile &T) {
    const int W = P.wmb * 16, H = P.hmb * 16, Wc = W >> 1;
    const int comp = lane < 16 ? 0 : lane < 24 ? 1 : 2;
    const int k = comp ? (lane & 7) : lane;
    int x0, y0;
    mb_origin(P, a, 0, x0, y0);
    const uint32_t v = P.bs[(size_t)a * 64 + lane], hraw = P.bs[(size_t)a * 64 + 32 + lane];
    const uint32_t h = (hraw & 0xF) | (((hraw >> 8) & 0xFFF) << 4);       // slots (0, 2, 3, 4) -> steps (0, 1, 2, 3)
    const H264B2MbInfo I = P.info[a];
    int qq = I.mb_class == H264B2_MB_IPCM ? 0 : I.qpy, qa = qq, qb = qq;
    if (A >= 0) { const H264B2MbInfo IA = P.info[A]; qa = IA.mb_class == H264B2_MB_IPCM ? 0 : IA.qpy; }
    if (B >= 0) { const H264B2MbInfo IB = P.info[B]; qb = IB.mb_class == H264B2_MB_IPCM ? 0 : IB.qpy; }
    if (comp) { qq = chroma_qp(P, qq, comp - 1); qa = chroma_qp(P, qa, comp - 1); qb = chroma_qp(P, qb, comp - 1); }
    DbThr tA, tB, tI;
    {
        const int oa = I.filter_offset_a, ob = I.filter_offset_b;
        const int q3[3] = { qa, qb, qq };
        DbThr *t3[3] = { &tA, &tB, &tI };
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int qpav = (q3[i] + qq + 1) >> 1;
            *t3[i] = db_thr_unpack(db_thr_pack(clip3i(0, 51, qpav + oa), clip3i(0, 51, qpav + ob)));
        }
    }
    const int topf = __any_sync(0xffffffffu, (h & 15u) != 0);
    uint8_t *Y = P.dst + (size_t)y0 * W + x0;
    uint8_t *Cbase = P.dst + (size_t)W * H;
    const size_t cplane = (size_t)Wc * (H >> 1);
    const int yc = y0 >> 1, xc = x0 >> 1;
    // stage: own samples, left columns, rows above
#pragma unroll
    for (int t = 0; t < 2; t++) { const int w = lane + 32 * t; T.L[4 + (w >> 2)][1 + (w & 3)] = __ldcg((const uint32_t *)(Y + (size_t)(w >> 2) * W + (w & 3) * 4)); }
    { const int c = lane >> 4, cl = lane & 15; T.C[c][4 + (cl >> 1)][1 + (cl & 1)] = __ldcg((const uint32_t *)(Cbase + c * cplane + (size_t)(yc + (cl >> 1)) * Wc + xc + (cl & 1) * 4)); }
    if (x0 > 0) {
        if (lane < 16) T.L[4 + lane][0] = __ldcg((const uint32_t *)(Y + (size_t)lane * W - 4));
        else { const int l = lane - 16, c = l >> 3, r = l & 7; T.C[c][4 + r][0] = __ldcg((const uint32_t *)(Cbase + c * cplane + (size_t)(yc + r) * Wc + xc - 4)); }
    }
    if (topf) {
        if (lane < 16) T.L[lane >> 2][1 + (lane & 3)] = __ldcg((const uint32_t *)(Y - (size_t)(4 - (lane >> 2)) * W + (lane & 3) * 4));
        else if (lane < 24) { const int l = lane - 16, c = l >> 2, r = (l >> 1) & 1, w = l & 1;
            T.C[c][2 + r][1 + w] = __ldcg((const uint32_t *)(Cbase + c * cplane + (size_t)(yc - 2 + r) * Wc + xc + w * 4)); }
    }
    __syncwarp();
    db_vertical(T, comp, k, v, tA, tI);
    __syncwarp();
    db_horizontal(T, comp, k, h, tB, tI);
    __syncwarp();
    // write back everything this MB may have changed
#pragma unroll
    for (int t = 0; t < 2; t++) { const int w = lane + 32 * t; *(uint32_t *)(Y + (size_t)(w >> 2) * W + (w & 3) * 4) = T.L[4 + (w >> 2)][1 + (w & 3)]; }
    { const int c = lane >> 4, cl = lane & 15; *(uint32_t *)(Cbase + c * cplane + (size_t)(yc + (cl >> 1)) * Wc + xc + (cl & 1) * 4) = T.C[c][4 + (cl >> 1)][1 + (cl & 1)]; }
    if (x0 > 0) {
        if (lane < 16) *(uint32_t *)(Y + (size_t)lane * W - 4) = T.L[4 + lane][0];
        else { const int l = lane - 16, c = l >> 3, r = l & 7; *(uint32_t *)(Cbase + c * cplane + (size_t)(yc + r) * Wc + xc - 4) = T.C[c][4 + r][0]; }
    }
    if (topf) {
        if (lane < 12) { const int rr = 1 + lane / 4, jj = lane & 3; *(uint32_t *)(Y + (size_t)(rr - 4) * W + jj * 4) = T.L[rr][1 + jj]; }
        else if (lane >= 16 && lane < 20) { const int l = lane - 16, cc = l >> 1, jj = l & 1; *(uint32_t *)(Cbase + cc * cplane + (size_t)(yc - 1) * Wc + xc + jj * 4) = T.C[cc][3][1 + jj]; }
    }
    __syncwarp();
}
