// CPU check of csrc/simd16.cuh (the packed two-line deblocking filter and the block re-formatting of k_deblock) against the
// scalar filter equations of the reference (H264PictureDeblockingFilterProcess.cpp:1314-1522).  Built and run by tests/test_simd_filter.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include "simd16.cuh"

static const uint8_t alpha_tab[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
static const uint8_t beta_tab[52]  = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18};
static const uint8_t tc0_tab[3][52] = {
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,4,5,6,6,7,8,9,10,11,13},
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,5,5,6,7,8,8,10,11,12,13,15,17},
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,4,5,6,6,7,8,9,10,11,13,14,16,18,20,23,25}};
static int clip3(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }
static int clip255(int v) { return clip3(0, 255, v); }

// one sample line, s[0..7] = p3 p2 p1 p0 q0 q1 q2 q3
static void scalar_line(uint8_t *s, int bS, int alpha, int beta, int ia, int chroma) {
    if (!bS) return;
    const int p3 = s[0], p2 = s[1], p1 = s[2], p0 = s[3], q0 = s[4], q1 = s[5], q2 = s[6], q3 = s[7];
    if (!(abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta)) return;
    int np0 = p0, np1 = p1, np2 = p2, nq0 = q0, nq1 = q1, nq2 = q2;
    const int ap = abs(p2 - p0), aq = abs(q2 - q0);
    if (bS < 4) {
        const int tc0 = tc0_tab[bS - 1][ia];
        const int tc = chroma ? tc0 + 1 : tc0 + (ap < beta) + (aq < beta);
        const int delta = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        np0 = clip255(p0 + delta); nq0 = clip255(q0 - delta);
        if (!chroma && ap < beta) np1 = p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1);
        if (!chroma && aq < beta) nq1 = q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1);
    } else {
        const int small = abs(p0 - q0) < ((alpha >> 2) + 2);
        if (!chroma && ap < beta && small) { np0 = (p2 + 2*p1 + 2*p0 + 2*q0 + q1 + 4) >> 3; np1 = (p2 + p1 + p0 + q0 + 2) >> 2; np2 = (2*p3 + 3*p2 + p1 + p0 + q0 + 4) >> 3; }
        else np0 = (2*p1 + p0 + q1 + 2) >> 2;
        if (!chroma && aq < beta && small) { nq0 = (p1 + 2*p0 + 2*q0 + 2*q1 + q2 + 4) >> 3; nq1 = (p0 + q0 + q1 + q2 + 2) >> 2; nq2 = (2*q3 + 3*q2 + q1 + q0 + p0 + 4) >> 3; }
        else nq0 = (2*q1 + q0 + p1 + 2) >> 2;
    }
    s[1] = (uint8_t)np2; s[2] = (uint8_t)np1; s[3] = (uint8_t)np0; s[4] = (uint8_t)nq0; s[5] = (uint8_t)nq1; s[6] = (uint8_t)nq2;
}

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main(int argc, char **argv) {
    const long iters = argc > 1 ? atol(argv[1]) : 2000000;
    long bad = 0, changed = 0;
    for (long it = 0; it < iters; it++) {
        // an 8x4 sample patch: rows 0..3, columns p3..q3; two 4x4 blocks P (cols 0..3) and Q (cols 4..7)
        uint8_t px[4][8];
        const int mode = rnd() % 4;
        const int base = rnd() & 255, spread = mode == 0 ? 256 : mode == 1 ? 4 : mode == 2 ? 16 : 64;
        for (int r = 0; r < 4; r++) for (int c = 0; c < 8; c++) px[r][c] = (uint8_t)clip255(mode == 0 ? (int)(rnd() & 255) : base + (int)(rnd() % spread) - spread / 2 + (c >= 4 ? (int)(rnd() % 24) - 12 : 0));
        const int ia = rnd() % 52, ib = rnd() % 52, chroma = rnd() & 1;
        const int alpha = alpha_tab[ia], beta = beta_tab[ib];
        int bS[4];                                   // per row
        for (int r = 0; r < 4; r++) bS[r] = rnd() % 5;
        const int vertical = rnd() & 1;              // 1: rows are the sample lines (vertical edge); 0: transpose roles (horizontal edge)
        // scalar result
        uint8_t ref[4][8];
        memcpy(ref, px, sizeof px);
        for (int r = 0; r < 4; r++) scalar_line(ref[r], bS[r], alpha, beta, ia, chroma);
        // packed result
        uint8_t got[4][8];
        DbPar2 k[2];
        for (int h = 0; h < 2; h++) {                // call h handles lines (h, h + 2)
            const int b0 = bS[h], b1 = bS[h + 2];
            k[h].kalpha = s16_ge_k(alpha); k[h].kbeta = s16_ge_k(beta); k[h].kalpha4 = s16_ge_k((alpha >> 2) + 2);
            k[h].tc0 = (uint32_t)((b0 >= 1 && b0 <= 3) ? tc0_tab[b0 - 1][ia] : 0) | ((uint32_t)((b1 >= 1 && b1 <= 3) ? tc0_tab[b1 - 1][ia] : 0) << 16);
            k[h].act = (b0 ? 0xFFFFu : 0u) | (b1 ? 0xFFFF0000u : 0u);
            k[h].s4 = (b0 == 4 ? 0xFFFFu : 0u) | (b1 == 4 ? 0xFFFF0000u : 0u);
            k[h].lum = chroma ? 0u : 0xFFFFFFFFu;
        }
        // warp-uniform path switches: exact (only the paths some line needs) or both always on
        const bool exact = rnd() & 1;
        bool do3 = !exact, do4 = !exact;
        for (int r = 0; r < 4; r++) { if (bS[r] >= 1 && bS[r] <= 3) do3 = true; if (bS[r] == 4) do4 = true; }
        if (vertical) {
            uint32_t rp[4], rq[4], pe[4], po[4], qe[4], qo[4];
            for (int r = 0; r < 4; r++) { memcpy(&rp[r], &px[r][0], 4); memcpy(&rq[r], &px[r][4], 4); }
            blk_rows_to_colpairs(rp, pe, po); blk_rows_to_colpairs(rq, qe, qo);
            db_filter2(pe[0], pe[1], pe[2], pe[3], qe[0], qe[1], qe[2], qe[3], k[0], do3, do4);
            db_filter2(po[0], po[1], po[2], po[3], qo[0], qo[1], qo[2], qo[3], k[1], do3, do4);
            blk_colpairs_to_rows(pe, po, rp); blk_colpairs_to_rows(qe, qo, rq);
            for (int r = 0; r < 4; r++) { memcpy(&got[r][0], &rp[r], 4); memcpy(&got[r][4], &rq[r], 4); }
        } else {
            // same data seen as a horizontal edge: sample line r of the patch = column r of a 4-wide strip, position c = row c
            uint32_t rp[4], rq[4], pe[4], po[4], qe[4], qo[4];
            for (int c = 0; c < 4; c++) { rp[c] = rq[c] = 0; for (int r = 0; r < 4; r++) { rp[c] |= (uint32_t)px[r][c] << (8 * r); rq[c] |= (uint32_t)px[r][4 + c] << (8 * r); } }
            blk_rows_to_rowpairs(rp, pe, po); blk_rows_to_rowpairs(rq, qe, qo);
            db_filter2(pe[0], pe[1], pe[2], pe[3], qe[0], qe[1], qe[2], qe[3], k[0], do3, do4);     // columns (0, 2) = lines 0, 2
            db_filter2(po[0], po[1], po[2], po[3], qo[0], qo[1], qo[2], qo[3], k[1], do3, do4);     // columns (1, 3) = lines 1, 3
            blk_rowpairs_to_rows(pe, po, rp); blk_rowpairs_to_rows(qe, qo, rq);
            for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) { got[r][c] = (uint8_t)(rp[c] >> (8 * r)); got[r][4 + c] = (uint8_t)(rq[c] >> (8 * r)); }
        }
        if (memcmp(got, ref, sizeof got)) {
            if (bad < 5) {
                fprintf(stderr, "mismatch it=%ld vertical=%d chroma=%d ia=%d ib=%d bS=%d,%d,%d,%d\n", it, vertical, chroma, ia, ib, bS[0], bS[1], bS[2], bS[3]);
                for (int r = 0; r < 4; r++) { fprintf(stderr, "  in "); for (int c = 0; c < 8; c++) fprintf(stderr, "%3d ", px[r][c]); fprintf(stderr, " ref "); for (int c = 0; c < 8; c++) fprintf(stderr, "%3d ", ref[r][c]); fprintf(stderr, " got "); for (int c = 0; c < 8; c++) fprintf(stderr, "%3d ", got[r][c]); fprintf(stderr, "\n"); }
            }
            bad++;
        }
        if (memcmp(ref, px, sizeof px)) changed++;
    }
    printf("iters %ld changed %ld bad %ld\n", iters, changed, bad);
    return bad != 0;
}
