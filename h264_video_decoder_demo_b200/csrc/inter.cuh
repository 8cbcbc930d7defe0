// inter.cuh — inter prediction fused with the inter residual add (T7/T8).
//
// Reference: Inter_prediction_process IP:412-667, fractional sample interpolation IP:2228-2328,
// luma 6-tap IP:2344-2480, chroma bilinear IP:2485-2522, weighted prediction IP:2526-2829,
// residual add IP:22-407, write-back IP:606-659 / PB:4408.
//
// v1 mapping: one CTA (256 threads) per macroblock, one thread per luma sample.  The motion field is
// the per-4x4 flattening of the reference's (mbPartIdx, subMbPartIdx) walk; every 4x4 block stages its
// own clamped 9x9 luma and 3x3 chroma reference windows in shared memory.
#pragma once
#include "common.cuh"
#include "residual.cuh"

struct RefViewDev {
    const uint8_t *base[3];
    int stride[3], wclamp[3], hclamp[3];
    int view;
};
// IP:2117 result code -> addressing (field views: PB:183-230; clamps IP:2351-2363, 2492-2512; Q4: a field
// view keeps the doubled stride as its width)
__device__ __forceinline__ void ref_view(const PicDev &P, int code, RefViewDev &rv) {
    const int slot = code >> 2, view = code & 3;
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
__device__ __forceinline__ int ref_px(const RefViewDev &rv, int c, int x, int y) {
    return __ldg(rv.base[c] + (size_t)clip3i(0, rv.hclamp[c] - 1, y) * rv.stride[c] + clip3i(0, rv.wclamp[c] - 1, x));
}
__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

// luma sample at (x,y) of a 4x4 block from its 9x9 window w (window origin = block origin - 2), IP:2344
__device__ inline int luma_interp_win(const uint8_t *w, int x, int y, int xF, int yF) {
#define S(dx, dy) ((int)w[(y + 2 + (dy)) * 9 + (x + 2 + (dx))])
    const int G = S(0,0);
    if (!xF && !yF) return G;
    const int b1 = tap6(S(-2,0), S(-1,0), G, S(1,0), S(2,0), S(3,0));
    const int h1 = tap6(S(0,-2), S(0,-1), G, S(0,1), S(0,2), S(0,3));
    const int b = clip255((b1 + 16) >> 5), h = clip255((h1 + 16) >> 5);
    if (!yF) { return xF == 2 ? b : xF == 1 ? (G + b + 1) >> 1 : (S(1,0) + b + 1) >> 1; }
    if (!xF) { return yF == 2 ? h : yF == 1 ? (G + h + 1) >> 1 : (S(0,1) + h + 1) >> 1; }
    const int s1 = tap6(S(-2,1), S(-1,1), S(0,1), S(1,1), S(2,1), S(3,1));
    const int m1 = tap6(S(1,-2), S(1,-1), S(1,0), S(1,1), S(1,2), S(1,3));
    const int s = clip255((s1 + 16) >> 5), m = clip255((m1 + 16) >> 5);
    int j = 0;
    if (xF == 2 || yF == 2) {
        const int cc = tap6(S(-2,-2), S(-2,-1), S(-2,0), S(-2,1), S(-2,2), S(-2,3));
        const int dd = tap6(S(-1,-2), S(-1,-1), S(-1,0), S(-1,1), S(-1,2), S(-1,3));
        const int ee = tap6(S(2,-2), S(2,-1), S(2,0), S(2,1), S(2,2), S(2,3));
        const int ff = tap6(S(3,-2), S(3,-1), S(3,0), S(3,1), S(3,2), S(3,3));
        j = clip255((tap6(cc, dd, h1, m1, ee, ff) + 512) >> 10);
    }
#undef S
    switch (xF * 4 + yF) {          // Table 8-12 (IP:2469)
    case 5:  return (b + h + 1) >> 1;   // e
    case 6:  return (h + j + 1) >> 1;   // i
    case 7:  return (h + s + 1) >> 1;   // p
    case 9:  return (b + j + 1) >> 1;   // f
    case 10: return j;
    case 11: return (j + s + 1) >> 1;   // q
    case 13: return (b + m + 1) >> 1;   // g
    case 14: return (j + m + 1) >> 1;   // k
    default: return (m + s + 1) >> 1;   // r (15)
    }
}

__device__ __forceinline__ int weigh(const H264B2Weight &w, int c, int have0, int have1, int p0, int p1) {   // IP:2617, IP:2699
    if (!w.mode) return (have0 && have1) ? (p0 + p1 + 1) >> 1 : have0 ? p0 : p1;
    const int ld = w.logwd[c];
    if (have0 && have1) return clip255(((p0 * w.w0[c] + p1 * w.w1[c] + (1 << ld)) >> (ld + 1)) + ((w.o0[c] + w.o1[c] + 1) >> 1));
    const int p = have0 ? p0 : p1, ww = have0 ? w.w0[c] : w.w1[c], oo = have0 ? w.o0[c] : w.o1[c];
    return ld >= 1 ? clip255(((p * ww + (1 << (ld - 1))) >> ld) + oo) : clip255(p * ww + oo);
}

struct InterSmem {
    ResidualTile rt;
    uint8_t lwin[2][16][84];      // [list][4x4 block][9x9]
    uint8_t cwin[2][2][16][12];   // [list][Cb/Cr][4x4 block][3x3]
};

__global__ void __launch_bounds__(256) k_inter(const PicDev *pics) {
    const PicDev &P = pics[blockIdx.y];
    const int a = blockIdx.x;
    if (a >= P.wmb * P.hmb) return;
    const H264B2MbInfo I = P.info[a];
    if (I.mb_class != H264B2_MB_INTER) return;
    __shared__ InterSmem sm;
    const int tid = threadIdx.x;
    const int field = P.mbaff && (I.flags & H264B2_MBF_FIELD);
    const int ys = field ? 2 : 1;
    int x0, y0;
    mb_origin(P, a, field, x0, y0);
    const int yA = field ? y0 / 2 : y0;                 // IP:577-580
    const int W = P.wmb * 16, H = P.hmb * 16, Wc = W >> 1;
    const H264B2MbMotion &M = P.motion[a];

    const int r = tid >> 4, px = tid & 15, x = px & 3, y = px >> 2;
    const int bx = (r & 3) * 4, by = (r >> 2) * 4, q = (by >> 3) * 2 + (bx >> 3);
    int have[2], mvx[2], mvy[2], mvcy[2];
#pragma unroll
    for (int l = 0; l < 2; l++) {
        const int code = M.ref_surf[l][q];
        have[l] = code >= 0;
        mvx[l] = M.mv[l][r][0]; mvy[l] = M.mv[l][r][1]; mvcy[l] = mvy[l];
        if (have[l]) {
            RefViewDev rv;
            ref_view(P, code, rv);
            if (field) { if (rv.view == 1 && (a & 1)) mvcy[l] += 2; else if (rv.view == 2 && !(a & 1)) mvcy[l] -= 2; }   // IP:2019-2043
            const int xI = x0 + bx + (mvx[l] >> 2) - 2, yI = yA + by + (mvy[l] >> 2) - 2;
            for (int i = px; i < 81; i += 16) sm.lwin[l][r][i] = (uint8_t)ref_px(rv, 0, xI + i % 9, yI + i / 9);
            const int xC = (x0 + bx) / 2 + (mvx[l] >> 3), yC = (yA + by) / 2 + (mvcy[l] >> 3);
            for (int i = px; i < 18; i += 16) { const int c = i / 9, j = i % 9; sm.cwin[l][c][r][j] = (uint8_t)ref_px(rv, 1 + c, xC + j % 3, yC + j / 3); }
        }
    }
    mb_residual(P, a, I, tid, 256, sm.rt, SyncCta());     // ends with __syncthreads(): windows are visible too
    if (!have[0] && !have[1]) return;                      // oracle: `continue` (nothing predicted, nothing added)
    const H264B2Weight w = P.weights[M.wt_idx[q]];
    // luma
    {
        int p[2] = {0, 0};
#pragma unroll
        for (int l = 0; l < 2; l++) if (have[l]) p[l] = luma_interp_win(sm.lwin[l][r], x, y, mvx[l] & 3, mvy[l] & 3);
        const int pred = weigh(w, 0, have[0], have[1], p[0], p[1]);
        const int yy = by + y, xx = bx + x;
        P.dst[(size_t)(y0 + yy * ys) * W + x0 + xx] = (uint8_t)clip255(pred + sm.rt.res[yy * 16 + xx]);
    }
    // chroma: threads 0..7 of each block: component c, 2x2 samples
    if (px < 8) {
        const int c = px >> 2, cx = px & 1, cy = (px >> 1) & 1;
        int p[2] = {0, 0};
#pragma unroll
        for (int l = 0; l < 2; l++) if (have[l]) {
            const uint8_t *cw = sm.cwin[l][c][r];
            const int xF = mvx[l] & 7, yF = mvcy[l] & 7;
            const int A = cw[cy * 3 + cx], B = cw[cy * 3 + cx + 1], C = cw[(cy + 1) * 3 + cx], D = cw[(cy + 1) * 3 + cx + 1];
            p[l] = ((8 - xF) * (8 - yF) * A + xF * (8 - yF) * B + (8 - xF) * yF * C + xF * yF * D + 32) >> 6;
        }
        const int pred = weigh(w, 1 + c, have[0], have[1], p[0], p[1]);
        const int yy = by / 2 + cy, xx = bx / 2 + cx;
        uint8_t *pl = P.dst + (size_t)W * H + (c ? (size_t)Wc * (H >> 1) : 0);
        pl[(size_t)(chroma_y0(y0) + yy * ys) * Wc + (x0 >> 1) + xx] = (uint8_t)clip255(pred + sm.rt.res[256 + c * 64 + yy * 8 + xx]);
    }
}
