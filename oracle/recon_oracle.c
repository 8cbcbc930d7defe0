/*
 * recon_oracle.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference decoder's picture-reconstruction path
 * (jfu222/h264_video_decoder_demo), consuming the same per-picture
 * structure-of-arrays as the B200 engine (include/h264_recon_b200.h).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product never routes through it.
 *
 * Parity is PINNED: tests/test_oracle_pin.py runs this code over the SoA that
 * oracle/ref_harness.cpp dumps from the unmodified reference for every
 * picture of the five bundled streams and requires the pre-deblock and
 * post-deblock picture checksums to equal the reference's own.
 *
 * Citations: PB = H264PictureBase.cpp, IP = H264InterPrediction.cpp,
 * DB = H264PictureDeblockingFilterProcess.cpp of the reference.
 * Macroblocks are processed strictly in address order like the reference
 * (H264SliceData.cpp:166-521), deblocking afterwards (PB:707).
 */
#include "h264_recon_b200.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
    const H264B2PicParams *p;
    int W, H, Wc, Hc, wmb, hmb, nmb, mbaff;
    uint8_t *plane[3];            /* destination Y, Cb, Cr */
    uint8_t *const *surf;         /* DPB surfaces by slot */
    int16_t ls4[2][2][6][16];     /* LevelScale4x4 in list order [inter][field scan][qP%6][k] (PB:4852) */
    int16_t ls8[2][2][6][64];
} Pic;

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* ------------------------------------------------------------------ scans (Table 8-13 / 8-14; PB:4542, PB:4597)
 * position = i*N + j for c[i][j] of the reference */
static const uint8_t zz4[16] = {0,1,4,8, 5,2,3,6, 9,12,13,10, 7,11,14,15};
static const uint8_t fs4[16] = {0,4,1,8, 12,5,9,13, 2,6,10,14, 3,7,11,15};
static uint8_t zz8[64];
static const uint8_t fs8[64] = {
     0, 8,16, 1, 9,24,32,17,  2,25,40,48,56,33,10, 3,
    18,41,49,57,26,11, 4,19, 34,42,50,58,27,12, 5,20,
    35,43,51,59,28,13, 6,21, 36,44,52,60,29,14,22,37,
    45,53,61,30, 7,15,38,46, 54,62,23,31,39,47,55,63 };
static int tables_ready = 0;
static void init_tables(void) {
    if (tables_ready) return;
    /* classic 8x8 zig-zag, first step to the right */
    int i = 0, j = 0;
    for (int k = 0; k < 64; k++) {
        zz8[k] = (uint8_t)(i * 8 + j);
        if ((i + j) % 2 == 0) { if (j == 7) i++; else if (i == 0) j++; else { i--; j++; } }
        else { if (i == 7) j++; else if (j == 0) i++; else { i++; j--; } }
    }
    tables_ready = 1;
}
static const int normadj4[6][3] = {{10,16,13},{11,18,14},{13,20,16},{14,23,18},{16,25,20},{18,29,23}};
static const int normadj8[6][6] = {{20,18,32,19,25,24},{22,19,35,21,28,26},{26,23,42,24,33,31},
                                   {28,25,45,26,35,33},{32,28,51,30,40,38},{36,32,58,34,46,43}};
static int na4(int m, int pos) { int i = pos >> 2, j = pos & 3; return (!(i & 1) && !(j & 1)) ? normadj4[m][0] : ((i & 1) && (j & 1)) ? normadj4[m][1] : normadj4[m][2]; }
static int na8(int m, int pos) {
    int i = pos >> 3, j = pos & 7;
    if (i % 4 == 0 && j % 4 == 0) return normadj8[m][0];
    if (i % 2 == 1 && j % 2 == 1) return normadj8[m][1];
    if (i % 4 == 2 && j % 4 == 2) return normadj8[m][2];
    if ((i % 4 == 0 && j % 2 == 1) || (i % 2 == 1 && j % 4 == 0)) return normadj8[m][3];
    if ((i % 4 == 0 && j % 4 == 2) || (i % 4 == 2 && j % 4 == 0)) return normadj8[m][4];
    return normadj8[m][5];
}
static void init_level_scale(Pic *P) {
    const H264B2PicParams *p = P->p;
    if (p->custom_scaling && p->level_scale4 && p->level_scale8) {
        memcpy(P->ls4, p->level_scale4, sizeof P->ls4); memcpy(P->ls8, p->level_scale8, sizeof P->ls8); return;
    }
    for (int inter = 0; inter < 2; inter++) for (int f = 0; f < 2; f++) for (int m = 0; m < 6; m++) {
        for (int k = 0; k < 16; k++) P->ls4[inter][f][m][k] = (int16_t)(16 * na4(m, f ? fs4[k] : zz4[k]));
        for (int k = 0; k < 64; k++) P->ls8[inter][f][m][k] = (int16_t)(16 * na8(m, f ? fs8[k] : zz8[k]));
    }
}

/* ------------------------------------------------------------------ geometry (PB:2503; MB.cpp:936) */
static inline int mb_field(const Pic *P, int a) { return (P->p->mb_info[a].flags & H264B2_MBF_FIELD) ? 1 : 0; }
static void mb_origin(const Pic *P, int a, int field, int *x0, int *y0) {
    if (!P->mbaff) { *x0 = (a % P->wmb) * 16; *y0 = (a / P->wmb) * 16; return; }
    int pair = a >> 1; *x0 = (pair % P->wmb) * 16; int yb = (pair / P->wmb) * 32;
    *y0 = field ? yb + (a & 1) : yb + (a & 1) * 16;
}
static inline int chroma_y0(int y0) { return (y0 >> 4) * 8 + (y0 & 1); }      /* PB:2133 */
/* luma 4x4 block index -> offset inside the MB (6.4.3) */
static inline int blk_x(int b) { return ((b >> 2) & 1) * 8 + (b & 1) * 4; }
static inline int blk_y(int b) { return (b >> 3) * 8 + ((b >> 1) & 1) * 4; }

/* ------------------------------------------------------------------ neighbouring locations (6.4.12; PB:2878, PB:2984)
 * returns neighbour MB address or -1; (xW,yW) inside that MB. */
static int avail_addr(const Pic *P, int cur, int n) {
    if (n < 0 || n > cur) return 0;
    return P->p->mb_info[n].slice_number == P->p->mb_info[cur].slice_number;
}
static int nbr_nonmbaff(const Pic *P, int cur, int xN, int yN, int maxW, int maxH, int *xW, int *yW) {
    int w = P->wmb, n = -1;
    if (xN < 0 && yN < 0)                        { n = cur - w - 1; if (cur % w == 0) n = -1; }
    else if (xN < 0 && yN < maxH)                { n = cur - 1;     if (cur % w == 0) n = -1; }
    else if (xN >= 0 && xN < maxW && yN < 0)     { n = cur - w; }
    else if (xN >= 0 && xN < maxW && yN >= 0 && yN < maxH) { *xW = xN; *yW = yN; return cur; }
    else if (xN >= maxW && yN < 0)               { n = cur - w + 1; if ((cur + 1) % w == 0) n = -1; }
    else return -1;
    if (n < 0 || !avail_addr(P, cur, n)) return -1;
    *xW = (xN + maxW) % maxW; *yW = (yN + maxH) % maxH;
    return n;
}
static int nbr_mbaff(const Pic *P, int cur, int xN, int yN, int maxW, int maxH, int *xW, int *yW) {
    const int w = P->wmb, pr = cur >> 1;
    int A = 2 * (pr - 1), B = 2 * (pr - w), C = 2 * (pr - w + 1), D = 2 * (pr - w - 1);
    if (!avail_addr(P, cur, A) || pr % w == 0) A = -2;
    if (!avail_addr(P, cur, B)) B = -2;
    if (!avail_addr(P, cur, C) || (pr + 1) % w == 0) C = -2;
    if (!avail_addr(P, cur, D) || pr % w == 0) D = -2;
    const int curFrame = !mb_field(P, cur), top = !(cur & 1);
    int n = -1, yM = 0;
#define XFRM(X) (!mb_field(P, (X)))
    if (xN < 0 && yN < 0) {
        if (curFrame) {
            if (top) { n = D + 1; yM = yN; }
            else if (A >= 0) { if (XFRM(A)) { n = A; yM = yN; } else { n = A + 1; yM = (yN + maxH) >> 1; } }   /* PB:3058-3077 */
        } else {
            if (top) { if (D >= 0) { if (XFRM(D)) { n = D + 1; yM = 2 * yN; } else { n = D; yM = yN; } } }
            else { n = D + 1; yM = yN; }
        }
    } else if (xN < 0 && yN >= 0 && yN < maxH) {
        if (A >= 0) {
            if (curFrame) {
                if (top) { if (XFRM(A)) { n = A; yM = yN; } else { n = (yN % 2 == 0) ? A : A + 1; yM = yN >> 1; } }
                else     { if (XFRM(A)) { n = A + 1; yM = yN; } else { n = (yN % 2 == 0) ? A : A + 1; yM = (yN + maxH) >> 1; } }
            } else {
                if (top) { if (XFRM(A)) { if (yN < maxH / 2) { n = A; yM = yN << 1; } else { n = A + 1; yM = (yN << 1) - maxH; } } else { n = A; yM = yN; } }
                else     { if (XFRM(A)) { if (yN < maxH / 2) { n = A; yM = (yN << 1) + 1; } else { n = A + 1; yM = (yN << 1) + 1 - maxH; } } else { n = A + 1; yM = yN; } }
            }
        }
    } else if (xN >= 0 && xN < maxW && yN < 0) {
        if (curFrame) { if (top) { n = B + 1; yM = yN; } else { n = cur - 1; yM = yN; } }
        else { if (top) { if (B >= 0) { if (XFRM(B)) { n = B + 1; yM = 2 * yN; } else { n = B; yM = yN; } } } else { n = B + 1; yM = yN; } }
    } else if (xN >= 0 && xN < maxW && yN >= 0 && yN < maxH) {
        *xW = xN; *yW = yN; return cur;
    } else if (xN >= maxW && yN < 0) {
        if (curFrame) { if (top) { n = C + 1; yM = yN; } else n = -1; }
        else { if (top) { if (C >= 0) { if (XFRM(C)) { n = C + 1; yM = 2 * yN; } else { n = C; yM = yN; } } } else { n = C + 1; yM = yN; } }
    }
#undef XFRM
    if (n < 0) return -1;
    *xW = (xN + maxW) % maxW; *yW = (yM + maxH) % maxH;
    return n;
}
static int nbr_loc(const Pic *P, int cur, int xN, int yN, int chroma, int *xW, int *yW) {
    int m = chroma ? 8 : 16;
    return P->mbaff ? nbr_mbaff(P, cur, xN, yN, m, m, xW, yW) : nbr_nonmbaff(P, cur, xN, yN, m, m, xW, yW);
}
/* constructed sample of the current picture at neighbouring location, or -1 (PB:1128-1154 etc.) */
static int nbr_sample(const Pic *P, int cur, int xN, int yN, int comp) {
    int xW, yW, n = nbr_loc(P, cur, xN, yN, comp != 0, &xW, &yW);
    if (n < 0) return -1;
    if (P->p->mb_info[n].flags & H264B2_MBF_CIP_UNAVAIL) return -1;
    int f = P->mbaff && mb_field(P, n), x0, y0;
    mb_origin(P, n, f, &x0, &y0);
    if (comp == 0) return P->plane[0][(y0 + (f ? 2 * yW : yW)) * P->W + x0 + xW];
    return P->plane[comp][(chroma_y0(y0) + (f ? 2 * yW : yW)) * P->Wc + (x0 >> 1) + xW];
}

/* ------------------------------------------------------------------ chroma QP (PB:4748) */
static int chroma_qp(const Pic *P, int qpy, int c /*0 Cb,1 Cr*/) {
    static const uint8_t tab[22] = {29,30,31,32,32,33,34,34,35,35,36,36,37,37,37,38,38,38,39,39,39,39};
    int qpi = clip3(0, 51, qpy + P->p->chroma_qp_offset[c]);
    return qpi < 30 ? qpi : tab[qpi - 30];
}

/* ------------------------------------------------------------------ residual transforms */
/* PB:4106: lv = 16 levels in list order; dc_pass: position (0,0) carries an already scaled DC */
static void resid4x4(const Pic *P, const int *lv, int qp, const int16_t (*ls)[16], int field, int dc_pass, int r[16]) {
    const uint8_t *scan = field ? fs4 : zz4;
    int d[16] = {0};
    for (int k = 0; k < 16; k++) {
        int pos = scan[k], c = lv[k];
        if (pos == 0 && dc_pass) d[0] = c;
        else if (qp >= 24) d[pos] = (c * ls[qp % 6][k]) << (qp / 6 - 4);
        else d[pos] = (c * ls[qp % 6][k] + (1 << (3 - qp / 6))) >> (4 - qp / 6);
    }
    (void)P;
    int f[16], h[16];
    for (int i = 0; i < 4; i++) {
        const int *s = d + 4 * i;
        int e0 = s[0] + s[2], e1 = s[0] - s[2], e2 = (s[1] >> 1) - s[3], e3 = s[1] + (s[3] >> 1);
        f[4*i] = e0 + e3; f[4*i+1] = e1 + e2; f[4*i+2] = e1 - e2; f[4*i+3] = e0 - e3;
    }
    for (int j = 0; j < 4; j++) {
        int g0 = f[j] + f[8+j], g1 = f[j] - f[8+j], g2 = (f[4+j] >> 1) - f[12+j], g3 = f[4+j] + (f[12+j] >> 1);
        h[j] = g0 + g3; h[4+j] = g1 + g2; h[8+j] = g1 - g2; h[12+j] = g0 - g3;
    }
    for (int i = 0; i < 16; i++) r[i] = (h[i] + 32) >> 6;
}
static void butterfly8(const int *in, int stride, int *out, int ostride) {   /* PB:4332-4390 one dimension */
    int a0 = in[0], a1 = in[stride], a2 = in[2*stride], a3 = in[3*stride], a4 = in[4*stride], a5 = in[5*stride], a6 = in[6*stride], a7 = in[7*stride];
    int e0 = a0 + a4, e1 = -a3 + a5 - a7 - (a7 >> 1), e2 = a0 - a4, e3 = a1 + a7 - a3 - (a3 >> 1);
    int e4 = (a2 >> 1) - a6, e5 = -a1 + a7 + a5 + (a5 >> 1), e6 = a2 + (a6 >> 1), e7 = a3 + a5 + a1 + (a1 >> 1);
    int f0 = e0 + e6, f1 = e1 + (e7 >> 2), f2 = e2 + e4, f3 = e3 + (e5 >> 2), f4 = e2 - e4, f5 = (e3 >> 2) - e5, f6 = e0 - e6, f7 = e7 - (e1 >> 2);
    out[0] = f0 + f7; out[ostride] = f2 + f5; out[2*ostride] = f4 + f3; out[3*ostride] = f6 + f1;
    out[4*ostride] = f6 - f1; out[5*ostride] = f4 - f3; out[6*ostride] = f2 - f5; out[7*ostride] = f0 - f7;
}
static void resid8x8(const int16_t *lv, int qp, const int16_t (*ls)[64], int field, int r[64]) {   /* PB:4270 */
    const uint8_t *scan = field ? fs8 : zz8;
    int d[64], g[64], m[64];
    memset(d, 0, sizeof d);
    for (int k = 0; k < 64; k++) {
        int c = lv[k];
        if (qp >= 36) d[scan[k]] = (c * ls[qp % 6][k]) << (qp / 6 - 6);
        else d[scan[k]] = (c * ls[qp % 6][k] + (1 << (5 - qp / 6))) >> (6 - qp / 6);
    }
    for (int i = 0; i < 8; i++) butterfly8(d + 8 * i, 1, g + 8 * i, 1);
    for (int j = 0; j < 8; j++) butterfly8(g + j, 8, m + j, 8);
    for (int i = 0; i < 64; i++) r[i] = (m[i] + 32) >> 6;
}
static void luma_dc16(const int16_t *lv, int qp, int ls00, int field, int dcY[16]) {   /* PB:4993; dcY raster [i][j] */
    const uint8_t *scan = field ? fs4 : zz4;
    int c[16] = {0}, g[16], f[16];
    for (int k = 0; k < 16; k++) c[scan[k]] = lv[k];
    for (int j = 0; j < 4; j++) {
        int a = c[j], b = c[4+j], cc = c[8+j], d = c[12+j];
        g[j] = a + b + cc + d; g[4+j] = a + b - cc - d; g[8+j] = a - b - cc + d; g[12+j] = a - b + cc - d;
    }
    for (int i = 0; i < 4; i++) {
        int a = g[4*i], b = g[4*i+1], cc = g[4*i+2], d = g[4*i+3];
        f[4*i] = a + b + cc + d; f[4*i+1] = a + b - cc - d; f[4*i+2] = a - b - cc + d; f[4*i+3] = a - b + cc - d;
    }
    for (int i = 0; i < 16; i++)
        dcY[i] = qp >= 36 ? (f[i] * ls00) << (qp / 6 - 6) : (f[i] * ls00 + (1 << (5 - qp / 6))) >> (6 - qp / 6);
}

/* walk the MB's coefficient blocks in storage order */
typedef struct { const int16_t *luma[16], *luma_dc, *chroma_dc, *cb[4], *cr[4], *pcm; } CoefPtrs;
static void coef_ptrs(const Pic *P, int a, CoefPtrs *cp) {
    const H264B2MbInfo *I = &P->p->mb_info[a];
    const int16_t *q = P->p->coefs + P->p->coef_offset[a];
    uint32_t m = I->coef_mask;
    memset(cp, 0, sizeof *cp);
    if (m & H264B2_CM_PCM) { cp->pcm = q; return; }
    int t8 = (I->flags & H264B2_MBF_T8x8) && I->mb_class != H264B2_MB_I16x16;
    for (int b = 0; b < 16; b++) if (m & H264B2_CM_LUMA(b)) { cp->luma[b] = q; q += t8 ? 64 : 16; }
    if (m & H264B2_CM_LUMA_DC) { cp->luma_dc = q; q += 16; }
    if (m & H264B2_CM_CHROMA_DC) { cp->chroma_dc = q; q += 8; }
    for (int b = 0; b < 4; b++) if (m & H264B2_CM_CB(b)) { cp->cb[b] = q; q += 16; }
    for (int b = 0; b < 4; b++) if (m & H264B2_CM_CR(b)) { cp->cr[b] = q; q += 16; }
}

/* add a residual block to the picture (clip), PB:3493 + PB:4408 */
static void add_block(uint8_t *pl, int stride, int x, int y, int ystep, const int *r, int n) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        uint8_t *px = &pl[(y + i * ystep) * stride + x + j];
        *px = (uint8_t)clip255(*px + r[i * n + j]);
    }
}

/* ------------------------------------------------------------------ intra prediction */
#define PA(x, y) pa[((y) + 1) * 17 + ((x) + 1)]
static void intra4x4_pred(const Pic *P, int a, int b, int mode, int x0, int y0, int ys) {   /* PB:1062 */
    int pa[5 * 17];
    const int xO = blk_x(b), yO = blk_y(b);
    for (int y = -1; y < 4; y++) PA(-1, y) = nbr_sample(P, a, xO - 1, yO + y, 0);
    for (int x = 0; x < 8; x++) PA(x, -1) = (x > 3 && (b == 3 || b == 11)) ? -1 : nbr_sample(P, a, xO + x, yO - 1, 0);
    if (PA(4,-1) < 0 && PA(5,-1) < 0 && PA(6,-1) < 0 && PA(7,-1) < 0 && PA(3,-1) >= 0) for (int x = 4; x < 8; x++) PA(x,-1) = PA(3,-1);
    int topok = PA(0,-1) >= 0 && PA(1,-1) >= 0 && PA(2,-1) >= 0 && PA(3,-1) >= 0;
    int trok = PA(4,-1) >= 0 && PA(5,-1) >= 0 && PA(6,-1) >= 0 && PA(7,-1) >= 0;
    int leftok = PA(-1,0) >= 0 && PA(-1,1) >= 0 && PA(-1,2) >= 0 && PA(-1,3) >= 0;
    int cornok = PA(-1,-1) >= 0;
    int pred[16], have = 0;
#define T(x) PA((x), -1)
#define L(y) PA(-1, (y))
    switch (mode) {
    case 0: if (topok) { have = 1; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) pred[y*4+x] = T(x); } break;
    case 1: if (leftok) { have = 1; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) pred[y*4+x] = L(y); } break;
    case 2: { int v; have = 1;
        if (topok && leftok) v = (T(0)+T(1)+T(2)+T(3)+L(0)+L(1)+L(2)+L(3)+4) >> 3;
        else if (leftok) v = (L(0)+L(1)+L(2)+L(3)+2) >> 2;
        else if (topok) v = (T(0)+T(1)+T(2)+T(3)+2) >> 2;
        else v = 128;
        for (int i = 0; i < 16; i++) pred[i] = v; } break;
    case 3: if (topok && trok) { have = 1; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
                pred[y*4+x] = (x == 3 && y == 3) ? (T(6) + 3*T(7) + 2) >> 2 : (T(x+y) + 2*T(x+y+1) + T(x+y+2) + 2) >> 2; } break;
    case 4: if (topok && leftok && cornok) { have = 1; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
                pred[y*4+x] = x > y ? (PA(x-y-2,-1) + 2*PA(x-y-1,-1) + PA(x-y,-1) + 2) >> 2
                            : x < y ? (PA(-1,y-x-2) + 2*PA(-1,y-x-1) + PA(-1,y-x) + 2) >> 2
                            : (T(0) + 2*PA(-1,-1) + L(0) + 2) >> 2; } break;
    case 5: if (topok && leftok && cornok) { have = 1; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
                int z = 2*x - y, v;
                if (z >= 0 && !(z & 1)) v = (PA(x-(y>>1)-1,-1) + PA(x-(y>>1),-1) + 1) >> 1;
                else if (z >= 0) v = (PA(x-(y>>1)-2,-1) + 2*PA(x-(y>>1)-1,-1) + PA(x-(y>>1),-1) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2*PA(-1,-1) + T(0) + 2) >> 2;
                else v = (PA(-1,y-1) + 2*PA(-1,y-2) + PA(-1,y-3) + 2) >> 2;
                pred[y*4+x] = v; } } break;
    case 6: if (topok && leftok && cornok) { have = 1; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
                int z = 2*y - x, v;
                if (z >= 0 && !(z & 1)) v = (PA(-1,y-(x>>1)-1) + PA(-1,y-(x>>1)) + 1) >> 1;
                else if (z >= 0) v = (PA(-1,y-(x>>1)-2) + 2*PA(-1,y-(x>>1)-1) + PA(-1,y-(x>>1)) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2*PA(-1,-1) + T(0) + 2) >> 2;
                else v = (PA(x-1,-1) + 2*PA(x-2,-1) + PA(x-3,-1) + 2) >> 2;
                pred[y*4+x] = v; } } break;
    case 7: if (topok && trok) { have = 1; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
                pred[y*4+x] = !(y & 1) ? (T(x+(y>>1)) + T(x+(y>>1)+1) + 1) >> 1 : (T(x+(y>>1)) + 2*T(x+(y>>1)+1) + T(x+(y>>1)+2) + 2) >> 2; } break;
    case 8: if (leftok) { have = 1; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
                int z = x + 2*y, v;
                if (z <= 4 && !(z & 1)) v = (L(y+(x>>1)) + L(y+(x>>1)+1) + 1) >> 1;
                else if (z < 5) v = (L(y+(x>>1)) + 2*L(y+(x>>1)+1) + L(y+(x>>1)+2) + 2) >> 2;
                else if (z == 5) v = (L(2) + 3*L(3) + 2) >> 2;
                else v = L(3);
                pred[y*4+x] = v; } } break;
    default: break;
    }
    if (have) for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) P->plane[0][(y0 + (yO + y) * ys) * P->W + x0 + xO + x] = (uint8_t)pred[y*4+x];
    /* otherwise the reference writes nothing: prediction = what the buffer holds (Q15) */
}

static void intra8x8_pred(const Pic *P, int a, int b, int mode, int x0, int y0, int ys) {   /* PB:1423 */
    int pa[9 * 17], q[9 * 17];
#define QA(x, y) q[((y) + 1) * 17 + ((x) + 1)]
    const int xO = (b & 1) * 8, yO = (b >> 1) * 8;
    for (int i = 0; i < 9 * 17; i++) { pa[i] = -1; q[i] = -1; }
    for (int y = -1; y < 8; y++) PA(-1, y) = nbr_sample(P, a, xO - 1, yO + y, 0);
    for (int x = 0; x < 16; x++) PA(x, -1) = nbr_sample(P, a, xO + x, yO - 1, 0);
    int trmiss = 1; for (int x = 8; x < 16; x++) if (PA(x,-1) >= 0) trmiss = 0;
    if (trmiss && PA(7,-1) >= 0) for (int x = 8; x < 16; x++) PA(x,-1) = PA(7,-1);
    int top16 = 1; for (int x = 0; x < 16; x++) if (PA(x,-1) < 0) top16 = 0;
    int left8 = 1; for (int y = 0; y < 8; y++) if (PA(-1,y) < 0) left8 = 0;
    /* reference sample filtering 8.3.2.2.1 (PB:1536-1599) */
    if (top16) {
        QA(0,-1) = PA(-1,-1) >= 0 ? (PA(-1,-1) + 2*PA(0,-1) + PA(1,-1) + 2) >> 2 : (3*PA(0,-1) + PA(1,-1) + 2) >> 2;
        for (int x = 1; x < 15; x++) QA(x,-1) = (PA(x-1,-1) + 2*PA(x,-1) + PA(x+1,-1) + 2) >> 2;
        QA(15,-1) = (PA(14,-1) + 3*PA(15,-1) + 2) >> 2;
    }
    if (PA(-1,-1) >= 0) {
        if (PA(0,-1) < 0 || PA(-1,0) < 0) {
            if (PA(0,-1) >= 0) QA(-1,-1) = (3*PA(-1,-1) + PA(0,-1) + 2) >> 2;
            else if (PA(-1,0) >= 0) QA(-1,-1) = (3*PA(-1,-1) + PA(-1,0) + 2) >> 2;
            else QA(-1,-1) = PA(-1,-1);
        } else QA(-1,-1) = (PA(0,-1) + 2*PA(-1,-1) + PA(-1,0) + 2) >> 2;
    }
    if (left8) {
        QA(-1,0) = PA(-1,-1) >= 0 ? (PA(-1,-1) + 2*PA(-1,0) + PA(-1,1) + 2) >> 2 : (3*PA(-1,0) + PA(-1,1) + 2) >> 2;
        for (int y = 1; y < 7; y++) QA(-1,y) = (PA(-1,y-1) + 2*PA(-1,y) + PA(-1,y+1) + 2) >> 2;
        QA(-1,7) = (PA(-1,6) + 3*PA(-1,7) + 2) >> 2;
    }
    int topok = 1; for (int x = 0; x < 8; x++) if (QA(x,-1) < 0) topok = 0;
    int trok = 1; for (int x = 8; x < 16; x++) if (QA(x,-1) < 0) trok = 0;
    int leftok = 1; for (int y = 0; y < 8; y++) if (QA(-1,y) < 0) leftok = 0;
    int cornok = QA(-1,-1) >= 0;
    int pred[64], have = 0;
    switch (mode) {
    case 0: if (topok) { have = 1; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) pred[y*8+x] = QA(x,-1); } break;
    case 1: if (leftok) { have = 1; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) pred[y*8+x] = QA(-1,y); } break;
    case 2: { int v = 0; have = 1;
        if (topok && leftok) { for (int i = 0; i < 8; i++) v += QA(i,-1) + QA(-1,i); v = (v + 8) >> 4; }
        else if (leftok) { for (int i = 0; i < 8; i++) v += QA(-1,i); v = (v + 4) >> 3; }
        else if (topok) { for (int i = 0; i < 8; i++) v += QA(i,-1); v = (v + 4) >> 3; }
        else v = 128;
        for (int i = 0; i < 64; i++) pred[i] = v; } break;
    case 3: if (topok && trok) { have = 1; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++)
                pred[y*8+x] = (x == 7 && y == 7) ? (QA(14,-1) + 3*QA(15,-1) + 2) >> 2 : (QA(x+y,-1) + 2*QA(x+y+1,-1) + QA(x+y+2,-1) + 2) >> 2; } break;
    case 4: if (topok && leftok && cornok) { have = 1; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++)
                pred[y*8+x] = x > y ? (QA(x-y-2,-1) + 2*QA(x-y-1,-1) + QA(x-y,-1) + 2) >> 2
                            : x < y ? (QA(-1,y-x-2) + 2*QA(-1,y-x-1) + QA(-1,y-x) + 2) >> 2
                            : (QA(0,-1) + 2*QA(-1,-1) + QA(-1,0) + 2) >> 2; } break;
    case 5: if (topok && leftok && cornok) { have = 1; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
                int z = 2*x - y, v;
                if (z >= 0 && !(z & 1)) v = (QA(x-(y>>1)-1,-1) + QA(x-(y>>1),-1) + 1) >> 1;
                else if (z >= 0) v = (QA(x-(y>>1)-2,-1) + 2*QA(x-(y>>1)-1,-1) + QA(x-(y>>1),-1) + 2) >> 2;
                else if (z == -1) v = (QA(-1,0) + 2*QA(-1,-1) + QA(0,-1) + 2) >> 2;
                else v = (QA(-1,y-2*x-1) + 2*QA(-1,y-2*x-2) + QA(-1,y-2*x-3) + 2) >> 2;
                pred[y*8+x] = v; } } break;
    case 6: if (topok && leftok && cornok) { have = 1; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
                int z = 2*y - x, v;
                if (z >= 0 && !(z & 1)) v = (QA(-1,y-(x>>1)-1) + QA(-1,y-(x>>1)) + 1) >> 1;
                else if (z >= 0) v = (QA(-1,y-(x>>1)-2) + 2*QA(-1,y-(x>>1)-1) + QA(-1,y-(x>>1)) + 2) >> 2;
                else if (z == -1) v = (QA(-1,0) + 2*QA(-1,-1) + QA(0,-1) + 2) >> 2;
                else v = (QA(x-2*y-1,-1) + 2*QA(x-2*y-2,-1) + QA(x-2*y-3,-1) + 2) >> 2;
                pred[y*8+x] = v; } } break;
    case 7: if (topok && trok) { have = 1; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++)
                pred[y*8+x] = !(y & 1) ? (QA(x+(y>>1),-1) + QA(x+(y>>1)+1,-1) + 1) >> 1 : (QA(x+(y>>1),-1) + 2*QA(x+(y>>1)+1,-1) + QA(x+(y>>1)+2,-1) + 2) >> 2; } break;
    case 8: if (leftok) { have = 1; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
                int z = x + 2*y, v;
                if (z <= 12 && !(z & 1)) v = (QA(-1,y+(x>>1)) + QA(-1,y+(x>>1)+1) + 1) >> 1;
                else if (z < 13) v = (QA(-1,y+(x>>1)) + 2*QA(-1,y+(x>>1)+1) + QA(-1,y+(x>>1)+2) + 2) >> 2;
                else if (z == 13) v = (QA(-1,6) + 3*QA(-1,7) + 2) >> 2;
                else v = QA(-1,7);
                pred[y*8+x] = v; } } break;
    default: break;
    }
    if (have) for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) P->plane[0][(y0 + (yO + y) * ys) * P->W + x0 + xO + x] = (uint8_t)pred[y*8+x];
#undef QA
}

static void intra16x16_pred(const Pic *P, int a, int mode, int x0, int y0, int ys) {   /* PB:1847 */
    int top[16], left[16], corner, topok = 1, leftok = 1;
    corner = nbr_sample(P, a, -1, -1, 0);
    for (int i = 0; i < 16; i++) { left[i] = nbr_sample(P, a, -1, i, 0); top[i] = nbr_sample(P, a, i, -1, 0); if (left[i] < 0) leftok = 0; if (top[i] < 0) topok = 0; }
    uint8_t *dst = P->plane[0]; const int W = P->W;
    if (mode == 0) { if (topok) for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) dst[(y0 + y*ys) * W + x0 + x] = (uint8_t)top[x]; }
    else if (mode == 1) { if (leftok) for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) dst[(y0 + y*ys) * W + x0 + x] = (uint8_t)left[y]; }
    else if (mode == 2) {
        int v = 0;
        if (topok && leftok) { for (int i = 0; i < 16; i++) v += top[i] + left[i]; v = (v + 16) >> 5; }
        else if (leftok) { for (int i = 0; i < 16; i++) v += left[i]; v = (v + 8) >> 4; }
        else if (topok) { for (int i = 0; i < 16; i++) v += top[i]; v = (v + 8) >> 4; }
        else v = 128;
        for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) dst[(y0 + y*ys) * W + x0 + x] = (uint8_t)v;
    } else if (topok && leftok) {             /* the reference does not test p[-1,-1] here (PB:2014-2018): -1 enters the sums */
        int Hh = 0, V = 0;
        for (int i = 0; i < 8; i++) { Hh += (i + 1) * (top[8 + i] - (6 - i >= 0 ? top[6 - i] : corner)); V += (i + 1) * (left[8 + i] - (6 - i >= 0 ? left[6 - i] : corner)); }
        int aa = 16 * (left[15] + top[15]), bb = (5 * Hh + 32) >> 6, cc = (5 * V + 32) >> 6;
        for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) dst[(y0 + y*ys) * W + x0 + x] = (uint8_t)clip255((aa + bb * (x - 7) + cc * (y - 7) + 16) >> 5);
    }
}

static void intra_chroma_pred(const Pic *P, int a, int comp, int mode, int xc0, int yc0, int ys) {   /* PB:2076 */
    int top[8], left[8], corner;
    corner = nbr_sample(P, a, -1, -1, comp);
    for (int i = 0; i < 8; i++) { left[i] = nbr_sample(P, a, -1, i, comp); top[i] = nbr_sample(P, a, i, -1, comp); }
    uint8_t *dst = P->plane[comp]; const int W = P->Wc;
    if (mode == 0) {
        for (int b = 0; b < 4; b++) {
            int xO = (b & 1) * 4, yO = (b >> 1) * 4, v;
            /* reference tests "> 0", so a neighbouring sample equal to 0 counts as unavailable (Q3, PB:2205-2250) */
            int t = top[xO] > 0 && top[xO+1] > 0 && top[xO+2] > 0 && top[xO+3] > 0;
            int l = left[yO] > 0 && left[yO+1] > 0 && left[yO+2] > 0 && left[yO+3] > 0;
            int st = top[xO] + top[xO+1] + top[xO+2] + top[xO+3], sl = left[yO] + left[yO+1] + left[yO+2] + left[yO+3];
            if ((xO == 0 && yO == 0) || (xO > 0 && yO > 0)) v = (t && l) ? (st + sl + 4) >> 3 : l ? (sl + 2) >> 2 : t ? (st + 2) >> 2 : 128;
            else if (xO > 0) v = t ? (st + 2) >> 2 : l ? (sl + 2) >> 2 : 128;
            else v = l ? (sl + 2) >> 2 : t ? (st + 2) >> 2 : 128;
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) dst[(yc0 + (yO + y) * ys) * W + xc0 + xO + x] = (uint8_t)v;
        }
        return;
    }
    int topok = 1, leftok = 1;
    for (int i = 0; i < 8; i++) { if (top[i] < 0) topok = 0; if (left[i] < 0) leftok = 0; }
    if (mode == 1) { if (leftok) for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) dst[(yc0 + y*ys) * W + xc0 + x] = (uint8_t)left[y]; }
    else if (mode == 2) { if (topok) for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) dst[(yc0 + y*ys) * W + xc0 + x] = (uint8_t)top[x]; }
    else if (topok && leftok && corner >= 0) {
        int Hh = 0, V = 0;
        for (int i = 0; i < 4; i++) { Hh += (i + 1) * (top[4 + i] - (2 - i >= 0 ? top[2 - i] : corner)); V += (i + 1) * (left[4 + i] - (2 - i >= 0 ? left[2 - i] : corner)); }
        int aa = 16 * (left[7] + top[7]), bb = (34 * Hh + 32) >> 6, cc = (34 * V + 32) >> 6;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) dst[(yc0 + y*ys) * W + xc0 + x] = (uint8_t)clip255((aa + bb * (x - 3) + cc * (y - 3) + 16) >> 5);
    }
}

/* ------------------------------------------------------------------ inter prediction */
typedef struct { const uint8_t *base[3]; int stride[3], wclamp[3], hclamp[3]; int view; } RefView;
static void ref_view(const Pic *P, int code, RefView *rv) {   /* IP:2117 result -> addressing (PB:183-230; IP:2351-2363, 2492-2512) */
    int slot = code >> 2, view = code & 3;
    const uint8_t *s = P->surf[slot];
    const uint8_t *pl[3] = { s, s + (size_t)P->W * P->H, s + (size_t)P->W * P->H + (size_t)P->Wc * P->Hc };
    int w[3] = { P->W, P->Wc, P->Wc }, h[3] = { P->H, P->Hc, P->Hc };
    rv->view = view;
    for (int c = 0; c < 3; c++) {
        if (view == 0) { rv->base[c] = pl[c]; rv->stride[c] = w[c]; rv->wclamp[c] = w[c]; rv->hclamp[c] = h[c]; }
        else { rv->base[c] = pl[c] + (view == 2 ? w[c] : 0); rv->stride[c] = 2 * w[c]; rv->wclamp[c] = 2 * w[c] /* Q4 */; rv->hclamp[c] = h[c] / 2; }
    }
}
static inline int ref_px(const RefView *rv, int c, int x, int y) {
    return rv->base[c][clip3(0, rv->hclamp[c] - 1, y) * rv->stride[c] + clip3(0, rv->wclamp[c] - 1, x)];
}
static inline int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }
static int luma_interp(const RefView *rv, int xI, int yI, int xF, int yF) {   /* IP:2344 */
#define S(dx, dy) ref_px(rv, 0, xI + (dx), yI + (dy))
    int G = S(0,0);
    if (!xF && !yF) return G;
    int b1 = tap6(S(-2,0), S(-1,0), G, S(1,0), S(2,0), S(3,0));
    int h1 = tap6(S(0,-2), S(0,-1), G, S(0,1), S(0,2), S(0,3));
    int b = clip255((b1 + 16) >> 5), h = clip255((h1 + 16) >> 5);
    int s1 = tap6(S(-2,1), S(-1,1), S(0,1), S(1,1), S(2,1), S(3,1));
    int m1 = tap6(S(1,-2), S(1,-1), S(1,0), S(1,1), S(1,2), S(1,3));
    int s = clip255((s1 + 16) >> 5), m = clip255((m1 + 16) >> 5);
    int cc = tap6(S(-2,-2), S(-2,-1), S(-2,0), S(-2,1), S(-2,2), S(-2,3));
    int dd = tap6(S(-1,-2), S(-1,-1), S(-1,0), S(-1,1), S(-1,2), S(-1,3));
    int ee = tap6(S(2,-2), S(2,-1), S(2,0), S(2,1), S(2,2), S(2,3));
    int ff = tap6(S(3,-2), S(3,-1), S(3,0), S(3,1), S(3,2), S(3,3));
    int j = clip255((tap6(cc, dd, h1, m1, ee, ff) + 512) >> 10);
    int H = S(1,0), M = S(0,1);
#undef S
    switch (xF * 4 + yF) {          /* Table 8-12 via predPartLXLs[xFrac][yFrac] (IP:2469) */
    case 1: return (G + h + 1) >> 1;   /* d */
    case 2: return h;
    case 3: return (M + h + 1) >> 1;   /* n */
    case 4: return (G + b + 1) >> 1;   /* a */
    case 5: return (b + h + 1) >> 1;   /* e */
    case 6: return (h + j + 1) >> 1;   /* i */
    case 7: return (h + s + 1) >> 1;   /* p */
    case 8: return b;
    case 9: return (b + j + 1) >> 1;   /* f */
    case 10: return j;
    case 11: return (j + s + 1) >> 1;  /* q */
    case 12: return (H + b + 1) >> 1;  /* c */
    case 13: return (b + m + 1) >> 1;  /* g */
    case 14: return (j + m + 1) >> 1;  /* k */
    default: return (m + s + 1) >> 1;  /* r */
    }
}
static int chroma_interp(const RefView *rv, int c, int xI, int yI, int xF, int yF) {   /* IP:2485 */
    int A = ref_px(rv, c, xI, yI), B = ref_px(rv, c, xI + 1, yI), C = ref_px(rv, c, xI, yI + 1), D = ref_px(rv, c, xI + 1, yI + 1);
    return ((8 - xF) * (8 - yF) * A + xF * (8 - yF) * B + (8 - xF) * yF * C + xF * yF * D + 32) >> 6;
}
static int weigh(const H264B2Weight *w, int c, int have0, int have1, int p0, int p1) {   /* IP:2617, IP:2699 */
    if (!w->mode) return (have0 && have1) ? (p0 + p1 + 1) >> 1 : have0 ? p0 : p1;
    int ld = w->logwd[c];
    if (have0 && have1) return clip255(((p0 * w->w0[c] + p1 * w->w1[c] + (1 << ld)) >> (ld + 1)) + ((w->o0[c] + w->o1[c] + 1) >> 1));
    int p = have0 ? p0 : p1, ww = have0 ? w->w0[c] : w->w1[c], oo = have0 ? w->o0[c] : w->o1[c];
    return ld >= 1 ? clip255(((p * ww + (1 << (ld - 1))) >> ld) + oo) : clip255(p * ww + oo);
}
static void inter_pred_mb(const Pic *P, int a, int x0, int y0, int ys, int field) {   /* IP:412 */
    const H264B2MbMotion *M = &P->p->motion[a];
    const int yA = field ? y0 / 2 : y0;         /* IP:577-580 */
    for (int r = 0; r < 16; r++) {
        int bx = (r & 3) * 4, by = (r >> 2) * 4, q = (by >> 3) * 2 + (bx >> 3);
        int have[2]; RefView rv[2];
        for (int l = 0; l < 2; l++) { have[l] = M->ref_surf[l][q] >= 0; if (have[l]) ref_view(P, M->ref_surf[l][q], &rv[l]); }
        if (!have[0] && !have[1]) continue;
        const H264B2Weight *w = &P->p->weights[M->wt_idx[q]];
        int pl[2][16], pc[2][2][4];
        for (int l = 0; l < 2; l++) if (have[l]) {
            int mvx = M->mv[l][r][0], mvy = M->mv[l][r][1];
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
                pl[l][y*4+x] = luma_interp(&rv[l], x0 + bx + (mvx >> 2) + x, yA + by + (mvy >> 2) + y, mvx & 3, mvy & 3);
            int mvcy = mvy;                     /* IP:2019-2043: cross-parity field reference */
            if (field) { if (rv[l].view == 1 && (a & 1)) mvcy += 2; else if (rv[l].view == 2 && !(a & 1)) mvcy -= 2; }
            for (int y = 0; y < 2; y++) for (int x = 0; x < 2; x++) for (int c = 0; c < 2; c++)
                pc[l][c][y*2+x] = chroma_interp(&rv[l], 1 + c, (x0 + bx) / 2 + (mvx >> 3) + x, (yA + by) / 2 + (mvcy >> 3) + y, mvx & 7, mvcy & 7);
        }
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
            P->plane[0][(y0 + (by + y) * ys) * P->W + x0 + bx + x] = (uint8_t)weigh(w, 0, have[0], have[1], pl[0][y*4+x], pl[1][y*4+x]);
        int yc0 = chroma_y0(y0);
        for (int c = 0; c < 2; c++) for (int y = 0; y < 2; y++) for (int x = 0; x < 2; x++)
            P->plane[1 + c][(yc0 + (by / 2 + y) * ys) * P->Wc + (x0 + bx) / 2 + x] = (uint8_t)weigh(w, 1 + c, have[0], have[1], pc[0][c][y*2+x], pc[1][c][y*2+x]);
    }
}

/* ------------------------------------------------------------------ one macroblock (H264SliceData.cpp:400-487) */
static void chroma_residual(const Pic *P, int a, const CoefPtrs *cp, int inter, int field, int xc0, int yc0, int ys) {   /* PB:3752, IP:230 */
    const H264B2MbInfo *I = &P->p->mb_info[a];
    const int16_t (*ls)[16] = P->ls4[inter][field];      /* luma list of the same MB (Q7) */
    for (int c = 0; c < 2; c++) {
        int qpc = chroma_qp(P, I->qpy, c);
        int dc[4] = {0, 0, 0, 0};
        if (cp->chroma_dc) {                             /* PB:3989 */
            const int16_t *s = cp->chroma_dc + 4 * c;
            int e00 = s[0] + s[2], e01 = s[1] + s[3], e10 = s[0] - s[2], e11 = s[1] - s[3];
            int f[4] = { e00 + e01, e00 - e01, e10 + e11, e10 - e11 };
            for (int i = 0; i < 4; i++) dc[i] = ((f[i] * ls[qpc % 6][0]) << (qpc / 6)) >> 5;
        }
        for (int b = 0; b < 4; b++) {
            const int16_t *ac = c ? cp->cr[b] : cp->cb[b];
            if (!ac && !dc[b]) continue;
            int lv[16], r[16];
            lv[0] = dc[b]; for (int k = 1; k < 16; k++) lv[k] = ac ? ac[k] : 0;
            resid4x4(P, lv, qpc, ls, field, 1, r);
            add_block(P->plane[1 + c], P->Wc, xc0 + (b & 1) * 4, yc0 + (b >> 1) * 4 * ys, ys, r, 4);
        }
    }
}
static void reconstruct_mb(const Pic *P, int a) {
    const H264B2MbInfo *I = &P->p->mb_info[a];
    if (I->mb_class == H264B2_MB_NA) return;
    const int field = P->mbaff && (I->flags & H264B2_MBF_FIELD);
    const int scanfield = (I->flags & H264B2_MBF_FIELD) ? 1 : 0;        /* field_pic_flag | mb_field_decoding_flag, PB:3419 */
    const int ys = field ? 2 : 1;
    int x0, y0; mb_origin(P, a, field, &x0, &y0);
    const int xc0 = x0 >> 1, yc0 = chroma_y0(y0);
    CoefPtrs cp; coef_ptrs(P, a, &cp);
    const int qp = I->qpy;
    if (I->mb_class == H264B2_MB_IPCM) {                                /* PB:2449 */
        for (int i = 0; i < 256; i++) P->plane[0][(y0 + ys * (i / 16)) * P->W + x0 + i % 16] = (uint8_t)cp.pcm[i];
        for (int i = 0; i < 64; i++) { P->plane[1][(yc0 + ys * (i / 8)) * P->Wc + xc0 + i % 8] = (uint8_t)cp.pcm[256 + i];
                                       P->plane[2][(yc0 + ys * (i / 8)) * P->Wc + xc0 + i % 8] = (uint8_t)cp.pcm[320 + i]; }
        return;
    }
    const int inter = I->mb_class == H264B2_MB_INTER;
    const int t8 = (I->flags & H264B2_MBF_T8x8) != 0;
    if (inter) inter_pred_mb(P, a, x0, y0, ys, field);
    if (I->mb_class == H264B2_MB_I16x16) {                              /* PB:3507 */
        intra16x16_pred(P, a, I->pred16_chroma & 3, x0, y0, ys);
        int dcY[16] = {0};
        if (cp.luma_dc) luma_dc16(cp.luma_dc, qp, P->ls4[0][scanfield][qp % 6][0], scanfield, dcY);
        for (int b = 0; b < 16; b++) {
            int dc = dcY[(blk_y(b) >> 2) * 4 + (blk_x(b) >> 2)];
            if (!cp.luma[b] && !dc) continue;
            int lv[16], r[16]; lv[0] = dc; for (int k = 1; k < 16; k++) lv[k] = cp.luma[b] ? cp.luma[b][k] : 0;
            resid4x4(P, lv, qp, P->ls4[0][scanfield], scanfield, 1, r);
            add_block(P->plane[0], P->W, x0 + blk_x(b), y0 + blk_y(b) * ys, ys, r, 4);
        }
    } else if (t8) {                                                    /* PB:3651, IP:130 */
        for (int b = 0; b < 4; b++) {
            if (I->mb_class == H264B2_MB_I8x8) intra8x8_pred(P, a, b, (int)((P->p->intra_modes[a] >> (4 * b)) & 15), x0, y0, ys);
            if (!cp.luma[b]) continue;
            int r[64]; resid8x8(cp.luma[b], qp, P->ls8[inter][scanfield], scanfield, r);
            add_block(P->plane[0], P->W, x0 + (b & 1) * 8, y0 + (b >> 1) * 8 * ys, ys, r, 8);
        }
    } else {                                                            /* PB:3401, IP:22 */
        for (int b = 0; b < 16; b++) {
            if (I->mb_class == H264B2_MB_I4x4) intra4x4_pred(P, a, b, (int)((P->p->intra_modes[a] >> (4 * b)) & 15), x0, y0, ys);
            if (!cp.luma[b]) continue;
            int lv[16], r[16]; for (int k = 0; k < 16; k++) lv[k] = cp.luma[b][k];
            resid4x4(P, lv, qp, P->ls4[inter][scanfield], scanfield, 0, r);
            add_block(P->plane[0], P->W, x0 + blk_x(b), y0 + blk_y(b) * ys, ys, r, 4);
        }
    }
    if (!inter) { int cm = (I->pred16_chroma >> 2) & 3; intra_chroma_pred(P, a, 1, cm, xc0, yc0, ys); }
    /* reference order: Cb pred+residual, then Cr pred+residual (H264SliceData.cpp:409-415); they are independent planes */
    if (!inter) { int cm = (I->pred16_chroma >> 2) & 3; intra_chroma_pred(P, a, 2, cm, xc0, yc0, ys); }
    chroma_residual(P, a, &cp, inter, scanfield, xc0, yc0, ys);
}

/* ------------------------------------------------------------------ deblocking (DB:76-1522) */
static const uint8_t alpha_tab[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
static const uint8_t beta_tab[52]  = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18};
static const uint8_t tc0_tab[3][52] = {
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,4,5,6,6,7,8,9,10,11,13},
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,5,5,6,7,8,8,10,11,12,13,15,17},
 {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,4,4,4,5,6,6,7,8,9,10,11,13,14,16,18,20,23,25}};

static inline int is_intra_mode(const H264B2MbInfo *I) { return I->mb_class >= H264B2_MB_I4x4 && I->mb_class <= H264B2_MB_I16x16; }   /* Q11 */

static int derive_bs(const Pic *P, int mbaff, int ap, int aq, int xp, int yp, int xq, int yq, int vertical) {   /* DB:994 */
    const H264B2MbInfo *Ip = &P->p->mb_info[ap], *Iq = &P->p->mb_info[aq];
    int fp = (Ip->flags & H264B2_MBF_FIELD) != 0, fq = (Iq->flags & H264B2_MBF_FIELD) != 0;
    int mixed = mbaff && ap != aq && fp != fq;
    int intra = is_intra_mode(Ip) || is_intra_mode(Iq);
    int spsi = ((Ip->flags | Iq->flags) & H264B2_MBF_SPSI) != 0;
    if (ap != aq) {
        if ((!fp && !fq && (intra || spsi)) || (mbaff && vertical && (intra || spsi))) return 4;
    }
    if ((!mixed && (intra || spsi)) || (mixed && !vertical && (intra || spsi))) return 3;
    int bp = 8 * (yp / 8) + 4 * (xp / 8) + 2 * ((yp % 8) / 4) + ((xp % 8) / 4);
    int bq = 8 * (yq / 8) + 4 * (xq / 8) + 2 * ((yq % 8) / 4) + ((xq % 8) / 4);
    if (((Ip->nnz_mask >> bp) & 1) || ((Iq->nnz_mask >> bq) & 1)) return 2;
    if (mixed) return 1;
    if (!P->p->motion) return 0;              /* no inter MB in the picture: only I_PCM can get here, all refs NULL, no MVs */
    const H264B2MbMotion *Mp = &P->p->motion[ap], *Mq = &P->p->motion[aq];
    int qp_ = (yp / 8) * 2 + xp / 8, qq_ = (yq / 8) * 2 + xq / 8, rp = (yp / 4) * 4 + xp / 4, rq = (yq / 4) * 4 + xq / 4;
    int r0p = Mp->ref_ident[0][qp_], r1p = Mp->ref_ident[1][qp_], r0q = Mq->ref_ident[0][qq_], r1q = Mq->ref_ident[1][qq_];
    int f0p = Mp->ref_surf[0][qp_] >= 0, f1p = Mp->ref_surf[1][qp_] >= 0, f0q = Mq->ref_surf[0][qq_] >= 0, f1q = Mq->ref_surf[1][qq_] >= 0;
    if (!(((r0p == r0q && r1p == r1q) || (r0p == r1q && r1p == r0q)) && (f0p + f1p) == (f0q + f1q))) return 1;
    const int lim = (mbaff && fq) ? 2 : 4;                                      /* DB:1157 */
    int m0px = Mp->mv[0][rp][0], m0py = Mp->mv[0][rp][1], m1px = Mp->mv[1][rp][0], m1py = Mp->mv[1][rp][1];
    int m0qx = Mq->mv[0][rq][0], m0qy = Mq->mv[0][rq][1], m1qx = Mq->mv[1][rq][0], m1qy = Mq->mv[1][rq][1];
#define FAR(ax, ay, bx, by) (iabs((ax) - (bx)) >= 4 || iabs((ay) - (by)) >= lim)
    if (f0p && !f1p && f0q && !f1q && FAR(m0px, m0py, m0qx, m0qy)) return 1;
    if (f0p && !f1p && !f0q && f1q && FAR(m0px, m0py, m1qx, m1qy)) return 1;
    if (!f0p && f1p && f0q && !f1q && FAR(m1px, m1py, m0qx, m0qy)) return 1;
    if (!f0p && f1p && !f0q && f1q && FAR(m1px, m1py, m1qx, m1qy)) return 1;
    if (f0p && f1p && r0p != r1p && f0q && f1q && ((r0q == r0p && r1q == r1p) || (r0q == r1p && r1q == r0p))) {
        if (r0q == r0p && (FAR(m0px, m0py, m0qx, m0qy) || FAR(m1px, m1py, m1qx, m1qy))) return 1;
        else if (r0q == r1p && (FAR(m1px, m1py, m0qx, m0qy) || FAR(m0px, m0py, m1qx, m1qy))) return 1;
    }
    if (f0p && f1p && r0p == r1p && f0q && f1q && r0q == r1q && r0q == r0p) {
        /* DB:1289-1298: A || (B && C) || D  (missing parentheses in the reference, Q5) */
        int A = FAR(m0px, m0py, m0qx, m0qy), B = FAR(m1px, m1py, m1qx, m1qy), C = FAR(m0px, m0py, m1qx, m1qy), D = FAR(m1px, m1py, m0qx, m0qy);
        if (A || (B && C) || D) return 1;
    }
#undef FAR
    return 0;
}

static void filter_line(const Pic *P, int comp, int bS, int ap, int aq, uint8_t *px, int step) {   /* DB:844 tail, DB:1314-1522 */
    /* px points at q0; p_i = px[-(i+1)*step], q_i = px[i*step] */
    int p[4], q[4];
    for (int i = 0; i < 4; i++) { q[i] = px[i * step]; p[i] = px[-(i + 1) * step]; }
    const H264B2MbInfo *Ip = &P->p->mb_info[ap], *Iq = &P->p->mb_info[aq];
    int qpp = Ip->mb_class == H264B2_MB_IPCM ? 0 : Ip->qpy, qpq = Iq->mb_class == H264B2_MB_IPCM ? 0 : Iq->qpy;
    if (comp) { qpp = chroma_qp(P, qpp, comp - 1); qpq = chroma_qp(P, qpq, comp - 1); }
    int qpav = (qpp + qpq + 1) >> 1;
    int ia = clip3(0, 51, qpav + Iq->filter_offset_a), ib = clip3(0, 51, qpav + Iq->filter_offset_b);
    int alpha = alpha_tab[ia], beta = beta_tab[ib];
    if (!(bS != 0 && iabs(p[0] - q[0]) < alpha && iabs(p[1] - p[0]) < beta && iabs(q[1] - q[0]) < beta)) return;
    int ap_ = iabs(p[2] - p[0]), aq_ = iabs(q[2] - q[0]);
    int np[3] = { p[0], p[1], p[2] }, nq[3] = { q[0], q[1], q[2] };
    if (bS < 4) {
        int tc0 = tc0_tab[bS - 1][ia];
        int tc = comp ? tc0 + 1 : tc0 + (ap_ < beta) + (aq_ < beta);
        int delta = clip3(-tc, tc, (((q[0] - p[0]) << 2) + (p[1] - q[1]) + 4) >> 3);
        np[0] = clip255(p[0] + delta); nq[0] = clip255(q[0] - delta);
        if (!comp && ap_ < beta) np[1] = p[1] + clip3(-tc0, tc0, (p[2] + ((p[0] + q[0] + 1) >> 1) - (p[1] << 1)) >> 1);
        if (!comp && aq_ < beta) nq[1] = q[1] + clip3(-tc0, tc0, (q[2] + ((p[0] + q[0] + 1) >> 1) - (q[1] << 1)) >> 1);
    } else {
        int small = iabs(p[0] - q[0]) < ((alpha >> 2) + 2);
        if (!comp && ap_ < beta && small) { np[0] = (p[2] + 2*p[1] + 2*p[0] + 2*q[0] + q[1] + 4) >> 3; np[1] = (p[2] + p[1] + p[0] + q[0] + 2) >> 2; np[2] = (2*p[3] + 3*p[2] + p[1] + p[0] + q[0] + 4) >> 3; }
        else np[0] = (2*p[1] + p[0] + q[1] + 2) >> 2;
        if (!comp && aq_ < beta && small) { nq[0] = (p[1] + 2*p[0] + 2*q[0] + 2*q[1] + q[2] + 4) >> 3; nq[1] = (p[0] + q[0] + q[1] + q[2] + 2) >> 2; nq[2] = (2*q[3] + 3*q[2] + q[1] + q[0] + p[0] + 4) >> 3; }
        else nq[0] = (2*q[1] + q[0] + p[1] + 2) >> 2;
    }
    for (int i = 0; i < 3; i++) { px[i * step] = (uint8_t)nq[i]; px[-(i + 1) * step] = (uint8_t)np[i]; }
}

/* DB:639: one edge of nE sample lines.  (xE,yE) fixed coordinate of the edge inside the MB in units of the component. */
static void filter_edge(const Pic *P, int mbaff, int a, int comp, int nbr, int vertical, int fieldmode, int leftflag, int e) {
    const int field = (P->p->mb_info[a].flags & H264B2_MBF_FIELD) != 0;
    int xI, yI; mb_origin(P, a, mbaff && field, &xI, &yI);
    const int stride = comp ? P->Wc : P->W, size = comp ? 8 : 16, dy = 1 + fieldmode;
    const int xP = comp ? xI / 2 : xI, yP = comp ? (yI + 1) / 2 : yI;
    uint8_t *pl = P->plane[comp];
    for (int k = 0; k < size; k++) {
        int xE = vertical ? e : k, yE = vertical ? k : e;
        int ap = a, xp, yp, xq, yq;
        uint8_t *q0; int step;
        if (vertical) {
            q0 = &pl[(yP + dy * yE) * stride + xP + xE]; step = 1;
            xp = xE - 1; if (xp < 0) { if (nbr >= 0) ap = leftflag ? nbr + (yE % 2) : nbr; xp += size; }
            yp = yE; xq = xE; yq = yE;
        } else {
            q0 = &pl[(yP + dy * yE - (yE % 2)) * stride + xP + xE]; step = dy * stride;
            xp = xE; yp = (yE - 1) - (yE % 2); if (yp < 0) { if (nbr >= 0) ap = nbr; yp += size; }
            xq = xE; yq = yE - (yE % 2);
        }
        int s = comp ? 2 : 1;
        int bS = derive_bs(P, mbaff, ap, a, (xp * s) & 255, (yp * s) & 255, (xq * s) & 255, (yq * s) & 255, vertical);
        filter_line(P, comp, bS, ap, a, q0, step);
    }
}

static void deblock_picture(const Pic *P) {   /* DB:76 */
    const H264B2MbInfo *info = P->p->mb_info;
    int stop = P->p->deblock_stop_mb; if (stop > P->nmb) stop = P->nmb;
    for (int a = 0; a < stop; a++) {
        const H264B2MbInfo *I = &info[a];
        const int mbaff = P->mbaff;           /* per-MB MbaffFrameFlag == picture flag for every decoded MB */
        const int field = (I->flags & H264B2_MBF_FIELD) != 0, t8 = (I->flags & H264B2_MBF_T8x8) != 0, idc = I->deblock_idc;
        int xW, yW;
        int A = nbr_loc(P, a, -1, 0, 0, &xW, &yW), B = nbr_loc(P, a, 0, -1, 0, &xW, &yW);
        const int fieldInFrame = mbaff && field;
        const int internal = idc != 1;
        int left = !((!mbaff && a % P->wmb == 0) || (mbaff && (a >> 1) % P->wmb == 0) || idc == 1 || (idc == 2 && A < 0));
        int top = !((!mbaff && a < P->wmb) || (mbaff && (a >> 1) < P->wmb && field) || (mbaff && (a >> 1) < P->wmb && !field && !(a & 1)) || idc == 1 || (idc == 2 && B < 0));
        int leftflag = mbaff && a >= 2 && !field && (info[a - 2].flags & H264B2_MBF_FIELD);
        int dbltop = mbaff && !(a & 1) && a >= 2 * P->wmb && !field && (info[a - 2 * P->wmb + 1].flags & H264B2_MBF_FIELD);
        /* luma */
        if (left) filter_edge(P, mbaff, a, 0, A, 1, fieldInFrame, leftflag, 0);
        if (internal) { if (!t8) filter_edge(P, mbaff, a, 0, A, 1, fieldInFrame, 0, 4); filter_edge(P, mbaff, a, 0, A, 1, fieldInFrame, 0, 8); if (!t8) filter_edge(P, mbaff, a, 0, A, 1, fieldInFrame, 0, 12); }
        if (top) {
            if (dbltop) { filter_edge(P, mbaff, a, 0, B - 1, 0, 1, 0, 0); filter_edge(P, mbaff, a, 0, B, 0, 1, 0, 1); }
            else filter_edge(P, mbaff, a, 0, B, 0, fieldInFrame, 0, 0);
        }
        if (internal) { if (!t8) filter_edge(P, mbaff, a, 0, B, 0, fieldInFrame, 0, 4); filter_edge(P, mbaff, a, 0, B, 0, fieldInFrame, 0, 8); if (!t8) filter_edge(P, mbaff, a, 0, B, 0, fieldInFrame, 0, 12); }
        /* chroma: Cb then Cr per edge (DB:348-610) */
        if (left) for (int c = 1; c <= 2; c++) filter_edge(P, mbaff, a, c, A, 1, fieldInFrame, leftflag, 0);
        if (internal) for (int c = 1; c <= 2; c++) filter_edge(P, mbaff, a, c, A, 1, fieldInFrame, 0, 4);
        if (top) {
            if (dbltop) { for (int c = 1; c <= 2; c++) filter_edge(P, mbaff, a, c, B, 0, 1, 0, 0); for (int c = 1; c <= 2; c++) filter_edge(P, mbaff, a, c, B, 0, 1, 0, 1); }   /* both passes use mbAddrB (DB:474-497) */
            else for (int c = 1; c <= 2; c++) filter_edge(P, mbaff, a, c, B, 0, fieldInFrame, 0, 0);
        }
        if (internal) for (int c = 1; c <= 2; c++) filter_edge(P, mbaff, a, c, B, 0, fieldInFrame, 0, 4);
    }
}

/* ------------------------------------------------------------------ entry points */
#define ORACLE_STAGE_RECON   1
#define ORACLE_STAGE_DEBLOCK 2

/* surfaces: array of n_surfaces host I420 buffers (Y|Cb|Cr contiguous).  stages: bit0 reconstruct, bit1 deblock. */
int oracle_reconstruct_picture(const H264B2PicParams *p, uint8_t *const *surfaces, int n_surfaces, int stages) {
    init_tables();
    if (!p || !surfaces || p->dst_surface < 0 || p->dst_surface >= n_surfaces) return -1;
    Pic P; memset(&P, 0, sizeof P);
    P.p = p; P.wmb = p->width_mbs; P.hmb = p->height_mbs; P.nmb = P.wmb * P.hmb; P.mbaff = p->mbaff_frame_flag;
    P.W = P.wmb * 16; P.H = P.hmb * 16; P.Wc = P.W / 2; P.Hc = P.H / 2; P.surf = surfaces;
    uint8_t *d = surfaces[p->dst_surface];
    P.plane[0] = d; P.plane[1] = d + (size_t)P.W * P.H; P.plane[2] = P.plane[1] + (size_t)P.Wc * P.Hc;
    init_level_scale(&P);
    if (stages & ORACLE_STAGE_RECON) {
        if (p->clear_surface) memset(d, 0, (size_t)P.W * P.H * 3 / 2);
        for (int a = 0; a < P.nmb; a++) reconstruct_mb(&P, a);
    }
    if ((stages & ORACLE_STAGE_DEBLOCK) && p->deblock_enable) deblock_picture(&P);
    return 0;
}

uint64_t oracle_checksum(const uint8_t *data, size_t bytes) {
    const uint64_t K = 0x9E3779B97F4A7C15ULL;
    uint64_t acc = 0; size_t nw = bytes / 4;
    for (size_t i = 0; i < nw; i++) { uint32_t w; memcpy(&w, data + 4 * i, 4); acc += ((uint64_t)w + 1) * ((2 * (uint64_t)i + 1) * K); }
    return acc;
}

/* YUV420P -> BGR24, the reference's integer BT.601 conversion (H264PictureBase.cpp:440-468; the FlipLines twin :471-498
 * writes row H-1-y).  C integer division truncates toward zero, exactly as in the reference. */
int oracle_convert_bgr24(const uint8_t *i420, int W, int H, uint8_t *bgr, int width_bytes, int flip) {
    if (!i420 || !bgr || W <= 0 || H <= 0 || width_bytes < 3 * W) return -1;
    const uint8_t *pu = i420 + (size_t)W * H, *pv = pu + (size_t)W * H / 4;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const int Y = i420[(size_t)y * W + x], U = pu[(size_t)(y / 2) * (W / 2) + x / 2], V = pv[(size_t)(y / 2) * (W / 2) + x / 2];
        const int b = (1164 * (Y - 16) + 2018 * (U - 128)) / 1000;
        const int g = (1164 * (Y - 16) - 813 * (V - 128) - 391 * (U - 128)) / 1000;
        const int r = (1164 * (Y - 16) + 1596 * (V - 128)) / 1000;
        uint8_t *o = bgr + (size_t)(flip ? H - 1 - y : y) * width_bytes + 3 * x;
        o[0] = (uint8_t)clip255(b); o[1] = (uint8_t)clip255(g); o[2] = (uint8_t)clip255(r);
    }
    return 0;
}
