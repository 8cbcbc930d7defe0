// inter.cuh — inter prediction + inter residual add (T7/T8), one THREAD per
// 4x4 luma block (and its 2x2 Cb / 2x2 Cr samples).
//
// Reference: Inter_prediction_process IP:412-667, fractional sample interpolation IP:2228-2328,
// luma 6-tap IP:2344-2480, chroma bilinear IP:2485-2522, weighted prediction IP:2526-2829,
// residual add IP:22-407 (u = Clip1(pred + r)), write-back IP:606-659 / PB:4408.
//
// Design (B200): the reference fetches 36 clamped samples and runs 13 six-tap filters PER SAMPLE
// (IP:2344).  Here each thread pulls the 9x9 window of its 4x4 block straight from the reference
// surface in HBM/L2 as 27 aligned 32-bit words (fast path, window inside the picture; a clamped byte
// path handles borders exactly like IP:2363), keeps it in registers, and evaluates the 6-tap filter with
// DP4A on packed bytes: 4 horizontal outputs = 8 DP4A; vertical outputs reuse the same routine on a
// PRMT-transposed window.  The 16 fractional positions are composed from three shared intermediates
// (horizontal half, vertical half, centre j), so divergent lanes execute a union of three phases, not
// 16 separate paths.  A warp covers one 4x4-block row of 8 consecutive macroblocks: its stores are
// 128 contiguous bytes per pixel row.
#pragma once
#include "common.cuh"
#include "residual_kernel.cuh"

struct RefViewDev {
    const uint8_t *base[3];
    int stride[3], wclamp[3], hclamp[3];
    int view;
};
// IP:2117 result code -> addressing (field views: PB:183-230; clamps IP:2351-2363, 2492-2512; Q4: a field
// view keeps the doubled stride as its width)
__device__ __forceinline__ void ref_view(const PicDev &P, int code, RefViewDev &rv) {
    const int slot = min(code >> 2, P.spp - 1), view = code & 3;
    const int W = P.wmb * 16, H = P.hmb * 16, Wc = W >> 1, Hc = H >> 1;
    const uint8_t *s = P.stream_base + (size_t)slot * P.frame_bytes;
    const uint8_t *pl[3] = { s, s + (size_t)W * H, s + (size_t)W * H + (size_t)Wc * Hc };
    const int w[3] = { W, Wc, Wc }, h[3] = { H, Hc, Hc };
    rv.view = view;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (view == 0) { rv.base[c] = pl[c]; rv.stride[c] = w[c]; rv.wclamp[c] = w[c]; rv.hclamp[c] = h[c]; }
        else { rv.base[c] = pl[c] + (view == 2 ? w[c] : 0); rv.stride[c] = 2 * w[c]; rv.wclamp[c] = 2 * w[c]; rv.hclamp[c] = h[c] / 2; }
    }
}

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {      // unsigned bytes x signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// 6-tap (1,-5,20,20,-5,1) at 4 consecutive positions of a 9-byte line (a0 = bytes 0..3, a1 = bytes 4..7, b8 = byte 8)
__device__ __forceinline__ void tap6x4(uint32_t a0, uint32_t a1, uint32_t b8, int &h0, int &h1, int &h2, int &h3) {
    h0 = dp4a_us(a0, 0x1414FB01u, dp4a_us(a1, 0x000001FBu, 0));
    h1 = dp4a_us(a0, 0x14FB0100u, dp4a_us(a1, 0x0001FB14u, 0));
    h2 = dp4a_us(a0, 0xFB010000u, dp4a_us(a1, 0x01FB1414u, 0));
    h3 = dp4a_us(a0, 0x01000000u, dp4a_us(a1, 0xFB1414FBu, (int)(b8 & 0xffu)));
}
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7F7F7F7Fu); }   // per byte (a+b+1)>>1
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) { return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24); }
__device__ __forceinline__ uint32_t pack4_half(int a, int b, int c, int d) {   // Clip1((x + 16) >> 5)
    return pack4(clip255((a + 16) >> 5), clip255((b + 16) >> 5), clip255((c + 16) >> 5), clip255((d + 16) >> 5));
}
// transpose a 4x4 byte tile: rows r0..r3 (byte x = column x) -> columns c0..c3 (byte y = row y)
__device__ __forceinline__ void transpose4x4(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3) {
    const uint32_t t0 = __byte_perm(r0, r1, 0x5140), t1 = __byte_perm(r2, r3, 0x5140);
    const uint32_t t2 = __byte_perm(r0, r1, 0x7362), t3 = __byte_perm(r2, r3, 0x7362);
    c0 = __byte_perm(t0, t1, 0x5410); c1 = __byte_perm(t0, t1, 0x7632);
    c2 = __byte_perm(t2, t3, 0x5410); c3 = __byte_perm(t2, t3, 0x7632);
}

// Predicted 4x4 luma block (4 packed rows) for one list: Luma_sample_interpolation_process (IP:2344) for all
// 16 samples at once.  (xI,yI): integer sample position of the block's top-left sample in the reference view.
__device__ __forceinline__ void luma_block_pred(const uint8_t *base, int stride, int wclamp, int hclamp, int wfast,
                                                int xI, int yI, int xF, int yF, uint32_t out[4]) {
    // window: rows wy = 0..8 <-> y = yI-2+wy, bytes wx = 0..8 <-> x = xI-2+wx; A0 = bytes 0..3, A1 = 4..7, A2 = byte 8
    uint32_t A0[9], A1[9], A2[9];
    const bool allrows = yF != 0;
    const bool fast = xI >= 2 && xI + 6 < wfast && yI >= 2 && yI + 6 < hclamp;
    if (fast) {
        const uint8_t *p = base + (size_t)(yI - 2) * stride + (xI - 2);
        const int o = (int)((uintptr_t)p & 3);
        const uint32_t *pw = (const uint32_t *)(p - o);
        const int sw = stride >> 2, sh = o * 8;
#pragma unroll
        for (int wy = 0; wy < 9; wy++) {
            if (allrows || (wy >= 2 && wy < 6)) {
                const uint32_t w0 = __ldg(pw + wy * sw), w1 = __ldg(pw + wy * sw + 1), w2 = __ldg(pw + wy * sw + 2);
                A0[wy] = __funnelshift_r(w0, w1, sh); A1[wy] = __funnelshift_r(w1, w2, sh); A2[wy] = w2 >> sh;
            } else { A0[wy] = A1[wy] = A2[wy] = 0; }
        }
    } else {
        int xs[9];
#pragma unroll
        for (int wx = 0; wx < 9; wx++) xs[wx] = clip3i(0, wclamp - 1, xI - 2 + wx);
#pragma unroll
        for (int wy = 0; wy < 9; wy++) {
            if (allrows || (wy >= 2 && wy < 6)) {
                const uint8_t *row = base + (size_t)clip3i(0, hclamp - 1, yI - 2 + wy) * stride;
                A0[wy] = pack4(__ldg(row + xs[0]), __ldg(row + xs[1]), __ldg(row + xs[2]), __ldg(row + xs[3]));
                A1[wy] = pack4(__ldg(row + xs[4]), __ldg(row + xs[5]), __ldg(row + xs[6]), __ldg(row + xs[7]));
                A2[wy] = __ldg(row + xs[8]);
            } else { A0[wy] = A1[wy] = A2[wy] = 0; }
        }
    }
    if (!xF && !yF) {
#pragma unroll
        for (int y = 0; y < 4; y++) out[y] = __byte_perm(A0[y + 2], A1[y + 2], 0x5432);
        return;
    }
    const bool needJ = (xF == 2 && yF != 0) || (yF == 2 && xF != 0);
    const bool needH = xF != 0 && yF != 2;              // horizontal half-sample rows b / s
    const bool needV = yF != 0 && xF != 2;              // vertical half-sample columns h / m
    uint32_t Hh[4] = {0, 0, 0, 0}, Vh[4] = {0, 0, 0, 0}, J[4] = {0, 0, 0, 0};
    if (xF != 0) {
        // unclipped horizontal 6-tap sums; all 9 rows when j is needed, else the 4 rows of b (yF != 3) or s (yF == 3)
        int hb[9][4];
        const int ry = 2 + (yF == 3);
#pragma unroll
        for (int wy = 0; wy < 9; wy++) {
            if (needJ || (wy >= ry && wy < ry + 4)) tap6x4(A0[wy], A1[wy], A2[wy], hb[wy][0], hb[wy][1], hb[wy][2], hb[wy][3]);
            else { hb[wy][0] = hb[wy][1] = hb[wy][2] = hb[wy][3] = 0; }
        }
        if (needH) {
#pragma unroll
            for (int y = 0; y < 4; y++) {
                const int r0 = yF == 3 ? hb[y + 3][0] : hb[y + 2][0], r1 = yF == 3 ? hb[y + 3][1] : hb[y + 2][1];
                const int r2 = yF == 3 ? hb[y + 3][2] : hb[y + 2][2], r3 = yF == 3 ? hb[y + 3][3] : hb[y + 2][3];
                Hh[y] = pack4_half(r0, r1, r2, r3);
            }
        }
        if (needJ) {
#pragma unroll
            for (int y = 0; y < 4; y++) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    const int t = (hb[y][x] + hb[y + 5][x]) - 5 * (hb[y + 1][x] + hb[y + 4][x]) + 20 * (hb[y + 2][x] + hb[y + 3][x]);
                    v[x] = clip255((t + 512) >> 10);
                }
                J[y] = pack4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    if (needV) {
        // vertical half samples of columns cx..cx+3 (cx = 2: h, cx = 3: m): transpose, then the same DP4A routine
        uint32_t R[9];
#pragma unroll
        for (int wy = 0; wy < 9; wy++) R[wy] = xF == 3 ? __byte_perm(A0[wy], A1[wy], 0x6543) : __byte_perm(A0[wy], A1[wy], 0x5432);
        uint32_t Ca[4], Cb[4];
        transpose4x4(R[0], R[1], R[2], R[3], Ca[0], Ca[1], Ca[2], Ca[3]);
        transpose4x4(R[4], R[5], R[6], R[7], Cb[0], Cb[1], Cb[2], Cb[3]);
        uint32_t V[4];      // V[x] = packed column x, byte y = row y
#pragma unroll
        for (int x = 0; x < 4; x++) {
            int v0, v1, v2, v3;
            tap6x4(Ca[x], Cb[x], R[8] >> (8 * x), v0, v1, v2, v3);
            V[x] = pack4_half(v0, v1, v2, v3);
        }
        transpose4x4(V[0], V[1], V[2], V[3], Vh[0], Vh[1], Vh[2], Vh[3]);
    }
#pragma unroll
    for (int y = 0; y < 4; y++) {
        uint32_t r;
        if (xF == 0) {                          // d, h, n
            const uint32_t G = yF == 3 ? __byte_perm(A0[y + 3], A1[y + 3], 0x5432) : __byte_perm(A0[y + 2], A1[y + 2], 0x5432);
            r = yF == 2 ? Vh[y] : avg4(G, Vh[y]);
        } else if (yF == 0) {                   // a, b, c
            const uint32_t G = xF == 3 ? __byte_perm(A0[y + 2], A1[y + 2], 0x6543) : __byte_perm(A0[y + 2], A1[y + 2], 0x5432);
            r = xF == 2 ? Hh[y] : avg4(G, Hh[y]);
        } else if (xF == 2 && yF == 2) r = J[y];
        else if (xF == 2) r = avg4(J[y], Hh[y]);            // f, q
        else if (yF == 2) r = avg4(J[y], Vh[y]);            // i, k
        else r = avg4(Hh[y], Vh[y]);                        // e, g, p, r
        out[y] = r;
    }
}

// 2x2 chroma samples of one plane for one list (IP:2485): returns p[0..3] = (0,0),(1,0),(0,1),(1,1)
__device__ __forceinline__ void chroma_block_pred(const uint8_t *base, int stride, int wclamp, int hclamp, int wfast,
                                                  int xC, int yC, int xF, int yF, int p[4]) {
    int s[3][3];
    if (xC >= 0 && xC + 2 < wfast && yC >= 0 && yC + 2 < hclamp) {
        const uint8_t *q = base + (size_t)yC * stride + xC;
        const int o = (int)((uintptr_t)q & 3);
        const uint32_t *pw = (const uint32_t *)(q - o);
        const int sw = stride >> 2, sh = o * 8;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const uint32_t w = __funnelshift_r(__ldg(pw + r * sw), __ldg(pw + r * sw + 1), sh);
            s[r][0] = w & 0xff; s[r][1] = (w >> 8) & 0xff; s[r][2] = (w >> 16) & 0xff;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const uint8_t *row = base + (size_t)clip3i(0, hclamp - 1, yC + r) * stride;
#pragma unroll
            for (int c = 0; c < 3; c++) s[r][c] = __ldg(row + clip3i(0, wclamp - 1, xC + c));
        }
    }
    const int w00 = (8 - xF) * (8 - yF), w10 = xF * (8 - yF), w01 = (8 - xF) * yF, w11 = xF * yF;
#pragma unroll
    for (int y = 0; y < 2; y++)
#pragma unroll
        for (int x = 0; x < 2; x++)
            p[y * 2 + x] = (w00 * s[y][x] + w10 * s[y][x + 1] + w01 * s[y + 1][x] + w11 * s[y + 1][x + 1] + 32) >> 6;
}

__device__ __forceinline__ int weigh(const H264B2Weight &w, int c, int have0, int have1, int p0, int p1) {   // IP:2617, IP:2699
    if (!w.mode) return (have0 && have1) ? (p0 + p1 + 1) >> 1 : have0 ? p0 : p1;
    const int ld = w.logwd[c];
    if (have0 && have1) return clip255(((p0 * w.w0[c] + p1 * w.w1[c] + (1 << ld)) >> (ld + 1)) + ((w.o0[c] + w.o1[c] + 1) >> 1));
    const int p = have0 ? p0 : p1, ww = have0 ? w.w0[c] : w.w1[c], oo = have0 ? w.o0[c] : w.o1[c];
    return ld >= 1 ? clip255(((p * ww + (1 << (ld - 1))) >> ld) + oo) : clip255(p * ww + oo);
}

// Generic path: ONE 4x4 luma block (raster slot r = by4 * 4 + bx4) of inter macroblock `a` with its 2x2 Cb / Cr samples.
// Used by k_inter (inter_quad.cuh) for macroblocks whose 8x8 quadrants carry more than one motion vector.
__device__ __noinline__ void inter_block_generic(const PicDev &P, int a, const H264B2MbInfo &I, int r4) {
    const int by4 = r4 >> 2, bx4 = r4 & 3;
    const int field = P.mbaff && (I.flags & H264B2_MBF_FIELD);
    const int ys = field ? 2 : 1;
    int x0, y0;
    mb_origin(P, a, field, x0, y0);
    const int yA = field ? y0 / 2 : y0;                 // IP:577-580
    const int W = P.wmb * 16, H = P.hmb * 16, Wc = W >> 1;
    const H264B2MbMotion &M = P.motion[a];
    const int r = by4 * 4 + bx4, bx = bx4 * 4, by = by4 * 4, q = (by4 >> 1) * 2 + (bx4 >> 1);

    // Both lists run through ONE copy of the interpolation code (the loop is deliberately not unrolled: the kernel
    // was instruction-cache bound, ncu stall_no_instruction was its top stall reason at 80 KB of SASS).
    uint32_t pl[2][4];
    int pc[2][2][4];
    int have[2];
#pragma unroll
    for (int i = 0; i < 4; i++) { pl[0][i] = pl[1][i] = 0; pc[0][0][i] = pc[0][1][i] = pc[1][0][i] = pc[1][1][i] = 0; }
    have[0] = M.ref_surf[0][q] >= 0; have[1] = M.ref_surf[1][q] >= 0;
#pragma unroll 1
    for (int l = 0; l < 2; l++) {
        const int code = M.ref_surf[l][q];
        if (code < 0) continue;
        RefViewDev rv;
        ref_view(P, code, rv);
        const int mvx = M.mv[l][r][0], mvy = M.mv[l][r][1];
        int mvcy = mvy;
        if (field) { if (rv.view == 1 && (a & 1)) mvcy += 2; else if (rv.view == 2 && !(a & 1)) mvcy -= 2; }   // IP:2019-2043
        uint32_t tl[4];
        int tc0[4], tc1[4];
        luma_block_pred(rv.base[0], rv.stride[0], rv.wclamp[0], rv.hclamp[0], W, x0 + bx + (mvx >> 2), yA + by + (mvy >> 2), mvx & 3, mvy & 3, tl);
        const int xC = (x0 + bx) / 2 + (mvx >> 3), yC = (yA + by) / 2 + (mvcy >> 3);
        chroma_block_pred(rv.base[1], rv.stride[1], rv.wclamp[1], rv.hclamp[1], Wc, xC, yC, mvx & 7, mvcy & 7, tc0);
        chroma_block_pred(rv.base[2], rv.stride[2], rv.wclamp[2], rv.hclamp[2], Wc, xC, yC, mvx & 7, mvcy & 7, tc1);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (l == 0) { pl[0][i] = tl[i]; pc[0][0][i] = tc0[i]; pc[0][1][i] = tc1[i]; }
            else        { pl[1][i] = tl[i]; pc[1][0][i] = tc0[i]; pc[1][1][i] = tc1[i]; }
        }
    }
    uint8_t *Y = P.dst + (size_t)(y0 + by * ys) * W + x0 + bx;
    uint8_t *C0 = P.dst + (size_t)W * H + (size_t)(chroma_y0(y0) + (by / 2) * ys) * Wc + (x0 >> 1) + bx / 2;
    uint8_t *C1 = C0 + (size_t)Wc * (H >> 1);
    const int none = !have[0] && !have[1];     // the reference predicts nothing: the residual lands on what the buffer holds
    const H264B2Weight w = P.weights[min((int)M.wt_idx[q], P.n_weights - 1)];
    const uint32_t m = I.coef_mask;
    const int t8 = (I.flags & H264B2_MBF_T8x8) != 0;
    const int16_t *res = P.res + (size_t)a * RES_MB_STRIDE;

    // ---- luma: combine lists, add residual, store 4 rows
    uint32_t o[4];
    if (none) {
#pragma unroll
        for (int y = 0; y < 4; y++) o[y] = *(const uint32_t *)(Y + (size_t)y * ys * W);
    } else if (!w.mode) {
#pragma unroll
        for (int y = 0; y < 4; y++) o[y] = (have[0] && have[1]) ? avg4(pl[0][y], pl[1][y]) : have[0] ? pl[0][y] : pl[1][y];
    } else {
#pragma unroll
        for (int y = 0; y < 4; y++) {
            int v[4];
#pragma unroll
            for (int x = 0; x < 4; x++) v[x] = weigh(w, 0, have[0], have[1], (pl[0][y] >> (8 * x)) & 0xff, (pl[1][y] >> (8 * x)) & 0xff);
            o[y] = pack4(v[0], v[1], v[2], v[3]);
        }
    }
    if (luma_slot_coded(m, H264B2_MB_INTER, t8, r)) {
        const uint4 ra = *(const uint4 *)(res + r * 16), rb = *(const uint4 *)(res + r * 16 + 8);
        const uint32_t rw[8] = { ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w };
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const int r0 = (int16_t)(rw[2 * y] & 0xffff), r1 = (int16_t)(rw[2 * y] >> 16), r2 = (int16_t)(rw[2 * y + 1] & 0xffff), r3 = (int16_t)(rw[2 * y + 1] >> 16);
            o[y] = pack4(clip255((int)(o[y] & 0xff) + r0), clip255((int)((o[y] >> 8) & 0xff) + r1), clip255((int)((o[y] >> 16) & 0xff) + r2), clip255((int)(o[y] >> 24) + r3));
        }
    }
#pragma unroll
    for (int y = 0; y < 4; y++) *(uint32_t *)(Y + (size_t)y * ys * W) = o[y];

    // ---- chroma: 2x2 per plane
#pragma unroll
    for (int c = 0; c < 2; c++) {
        uint8_t *C = c ? C1 : C0;
        int v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = none ? C[(size_t)(i >> 1) * ys * Wc + (i & 1)] : weigh(w, 1 + c, have[0], have[1], pc[0][c][i], pc[1][c][i]);
        const int cb = (by4 >> 1) * 2 + (bx4 >> 1);
        if (chroma_blk_coded(m, c, cb)) {
            const int16_t *rc = res + (16 + c * 4 + cb) * 16 + ((by4 & 1) * 2) * 4 + (bx4 & 1) * 2;
            const uint32_t e0 = *(const uint32_t *)rc, e1 = *(const uint32_t *)(rc + 4);
            v[0] = clip255(v[0] + (int16_t)(e0 & 0xffff)); v[1] = clip255(v[1] + (int16_t)(e0 >> 16));
            v[2] = clip255(v[2] + (int16_t)(e1 & 0xffff)); v[3] = clip255(v[3] + (int16_t)(e1 >> 16));
        }
        *(uint16_t *)(C) = (uint16_t)(v[0] | (v[1] << 8));
        *(uint16_t *)(C + (size_t)ys * Wc) = (uint16_t)(v[2] | (v[3] << 8));
    }
}
