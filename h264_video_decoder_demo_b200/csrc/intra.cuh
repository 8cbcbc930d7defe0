// intra.cuh — intra prediction fused with the intra residual add, run as a
// macroblock wavefront: one warp owns one MB row (MB-pair row under MBAFF) and advances left to right;
// before an intra MB at column x it waits until the row above has published progress >= x+2, which
// covers the left (same warp), top, top-left and top-right neighbours.
//
// Reference: Intra_4x4 PB:1062-1419, Intra_8x8 PB:1423-1843 (reference sample filter PB:1536-1599),
// Intra_16x16 PB:1847-2056, chroma PB:2076-2381 (DC availability uses `> 0`, Q3), I_PCM PB:2449-2499,
// drivers PB:3401-3929 (block n predicts from block n-1's reconstruction).
#pragma once
#include "common.cuh"
#include "residual.cuh"
#include "residual_kernel.cuh"
#include "wavefront.cuh"

struct IntraWarpSmem {
    ResidualTile rt;
    int nb[36];
    int qa[28];
    // fast (progressive) path: the macroblock and its neighbour samples staged in shared memory.
    // Yt row 0 = samples of the row above (x = -1..23 at byte x+4), rows 1..16 = MB rows (byte 3 = left neighbour,
    // bytes 4..19 = the MB's own samples).  Ct: same for Cb / Cr (x = -1..7 at byte x+4).
    uint32_t Yt[17][8];
    uint32_t Ct[2][9][4];
};

// sample of the current picture at MB-relative luma/chroma location (xN,yN) from the staged tile, or -1 when the
// location is not available (6.4.12 for non-MBAFF frames, PB:2878; availability flags already include slice
// membership and constrained_intra_pred)
__device__ __forceinline__ int tile_sample(const IntraWarpSmem &S, int xN, int yN, int comp, int avA, int avB, int avC, int avD) {
    const int n = comp ? 8 : 16;
    if (yN < 0) {
        const int ok = xN < 0 ? avD : xN < n ? avB : avC;
        if (!ok) return -1;
        return comp ? ((const uint8_t *)S.Ct[comp - 1][0])[xN + 4] : ((const uint8_t *)S.Yt[0])[xN + 4];
    }
    if (xN < 0) { if (!avA) return -1; return comp ? ((const uint8_t *)S.Ct[comp - 1][yN + 1])[3] : ((const uint8_t *)S.Yt[yN + 1])[3]; }
    if (xN >= n) return -1;
    return comp ? ((const uint8_t *)S.Ct[comp - 1][yN + 1])[xN + 4] : ((const uint8_t *)S.Yt[yN + 1])[xN + 4];
}

__device__ __forceinline__ int ld_acquire_flag(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_flag(int *p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ---- 4x4 (nb: [0] corner, [1..4] left y=0..3, [5..12] top x=0..7) ----
__device__ inline int pred4x4_px(int mode, int x, int y, const int *nb, int &have) {
#define T(i) ((i) < 0 ? nb[0] : nb[5 + (i)])
#define L(i) ((i) < 0 ? nb[0] : nb[1 + (i)])
    const int topok = nb[5] >= 0 && nb[6] >= 0 && nb[7] >= 0 && nb[8] >= 0;
    const int trok = nb[9] >= 0 && nb[10] >= 0 && nb[11] >= 0 && nb[12] >= 0;
    const int leftok = nb[1] >= 0 && nb[2] >= 0 && nb[3] >= 0 && nb[4] >= 0;
    const int cornok = nb[0] >= 0;
    have = 0;
    switch (mode) {
    case 0: if (topok) { have = 1; return T(x); } break;
    case 1: if (leftok) { have = 1; return L(y); } break;
    case 2: have = 1;
        if (topok && leftok) return (T(0)+T(1)+T(2)+T(3)+L(0)+L(1)+L(2)+L(3)+4) >> 3;
        if (leftok) return (L(0)+L(1)+L(2)+L(3)+2) >> 2;
        if (topok) return (T(0)+T(1)+T(2)+T(3)+2) >> 2;
        return 128;
    case 3: if (topok && trok) { have = 1; return (x == 3 && y == 3) ? (T(6) + 3*T(7) + 2) >> 2 : (T(x+y) + 2*T(x+y+1) + T(x+y+2) + 2) >> 2; } break;
    case 4: if (topok && leftok && cornok) { have = 1;
            return x > y ? (T(x-y-2) + 2*T(x-y-1) + T(x-y) + 2) >> 2
                 : x < y ? (L(y-x-2) + 2*L(y-x-1) + L(y-x) + 2) >> 2
                 : (T(0) + 2*nb[0] + L(0) + 2) >> 2; } break;
    case 5: if (topok && leftok && cornok) { have = 1;
            const int z = 2*x - y;
            if (z >= 0 && !(z & 1)) return (T(x-(y>>1)-1) + T(x-(y>>1)) + 1) >> 1;
            if (z >= 0) return (T(x-(y>>1)-2) + 2*T(x-(y>>1)-1) + T(x-(y>>1)) + 2) >> 2;
            if (z == -1) return (L(0) + 2*nb[0] + T(0) + 2) >> 2;
            return (L(y-1) + 2*L(y-2) + L(y-3) + 2) >> 2; } break;
    case 6: if (topok && leftok && cornok) { have = 1;
            const int z = 2*y - x;
            if (z >= 0 && !(z & 1)) return (L(y-(x>>1)-1) + L(y-(x>>1)) + 1) >> 1;
            if (z >= 0) return (L(y-(x>>1)-2) + 2*L(y-(x>>1)-1) + L(y-(x>>1)) + 2) >> 2;
            if (z == -1) return (L(0) + 2*nb[0] + T(0) + 2) >> 2;
            return (T(x-1) + 2*T(x-2) + T(x-3) + 2) >> 2; } break;
    case 7: if (topok && trok) { have = 1;
            return !(y & 1) ? (T(x+(y>>1)) + T(x+(y>>1)+1) + 1) >> 1 : (T(x+(y>>1)) + 2*T(x+(y>>1)+1) + T(x+(y>>1)+2) + 2) >> 2; } break;
    case 8: if (leftok) { have = 1;
            const int z = x + 2*y;
            if (z <= 4 && !(z & 1)) return (L(y+(x>>1)) + L(y+(x>>1)+1) + 1) >> 1;
            if (z < 5) return (L(y+(x>>1)) + 2*L(y+(x>>1)+1) + L(y+(x>>1)+2) + 2) >> 2;
            if (z == 5) return (L(2) + 3*L(3) + 2) >> 2;
            return L(3); } break;
    default: break;
    }
#undef T
#undef L
    return 0;
}

// ---- 8x8 (qa: filtered samples, [0] corner, [1..8] left y=0..7, [9..24] top x=0..15) ----
__device__ inline int pred8x8_px(int mode, int x, int y, const int *qa, int topok, int trok, int leftok, int cornok, int &have) {
#define T(i) ((i) < 0 ? qa[0] : qa[9 + (i)])
#define L(i) ((i) < 0 ? qa[0] : qa[1 + (i)])
    have = 0;
    switch (mode) {
    case 0: if (topok) { have = 1; return T(x); } break;
    case 1: if (leftok) { have = 1; return L(y); } break;
    case 2: { have = 1; int v = 0;
        if (topok && leftok) { for (int i = 0; i < 8; i++) v += T(i) + L(i); return (v + 8) >> 4; }
        if (leftok) { for (int i = 0; i < 8; i++) v += L(i); return (v + 4) >> 3; }
        if (topok) { for (int i = 0; i < 8; i++) v += T(i); return (v + 4) >> 3; }
        return 128; }
    case 3: if (topok && trok) { have = 1; return (x == 7 && y == 7) ? (T(14) + 3*T(15) + 2) >> 2 : (T(x+y) + 2*T(x+y+1) + T(x+y+2) + 2) >> 2; } break;
    case 4: if (topok && leftok && cornok) { have = 1;
            return x > y ? (T(x-y-2) + 2*T(x-y-1) + T(x-y) + 2) >> 2
                 : x < y ? (L(y-x-2) + 2*L(y-x-1) + L(y-x) + 2) >> 2
                 : (T(0) + 2*qa[0] + L(0) + 2) >> 2; } break;
    case 5: if (topok && leftok && cornok) { have = 1;
            const int z = 2*x - y;
            if (z >= 0 && !(z & 1)) return (T(x-(y>>1)-1) + T(x-(y>>1)) + 1) >> 1;
            if (z >= 0) return (T(x-(y>>1)-2) + 2*T(x-(y>>1)-1) + T(x-(y>>1)) + 2) >> 2;
            if (z == -1) return (L(0) + 2*qa[0] + T(0) + 2) >> 2;
            return (L(y-2*x-1) + 2*L(y-2*x-2) + L(y-2*x-3) + 2) >> 2; } break;
    case 6: if (topok && leftok && cornok) { have = 1;
            const int z = 2*y - x;
            if (z >= 0 && !(z & 1)) return (L(y-(x>>1)-1) + L(y-(x>>1)) + 1) >> 1;
            if (z >= 0) return (L(y-(x>>1)-2) + 2*L(y-(x>>1)-1) + L(y-(x>>1)) + 2) >> 2;
            if (z == -1) return (L(0) + 2*qa[0] + T(0) + 2) >> 2;
            return (T(x-2*y-1) + 2*T(x-2*y-2) + T(x-2*y-3) + 2) >> 2; } break;
    case 7: if (topok && trok) { have = 1;
            return !(y & 1) ? (T(x+(y>>1)) + T(x+(y>>1)+1) + 1) >> 1 : (T(x+(y>>1)) + 2*T(x+(y>>1)+1) + T(x+(y>>1)+2) + 2) >> 2; } break;
    case 8: if (leftok) { have = 1;
            const int z = x + 2*y;
            if (z <= 12 && !(z & 1)) return (L(y+(x>>1)) + L(y+(x>>1)+1) + 1) >> 1;
            if (z < 13) return (L(y+(x>>1)) + 2*L(y+(x>>1)+1) + L(y+(x>>1)+2) + 2) >> 2;
            if (z == 13) return (L(6) + 3*L(7) + 2) >> 2;
            return L(7); } break;
    default: break;
    }
#undef T
#undef L
    return 0;
}

// Directional modes 3..8 as data.  Every predicted sample of those modes is (w0*P[i0] + w1*P[i1] + w2*P[i2] + 2) >> 2 over the
// block's neighbour array P (4x4: nb[0..12]; 8x8: the filtered qa[0..24]) with (w0,w1,w2) = (1,2,1), (2,2,0) — the two-tap
// average (a+b+1)>>1; a copy is (1,2,1) over one index.  entry = i0 | i1 << 5 | i2 << 10 | two_tap << 15; the host fills the tables at context
// creation from the same rules as pred4x4_px / pred8x8_px above (fill_intra_tables in engine.cu), and the kernels keep a copy
// in shared memory: no per-sample switch, no divergence between the samples of a block.
__device__ uint16_t g_pred4_tab[6][16];
__device__ uint16_t g_pred8_tab[6][64];
__device__ __forceinline__ int pred_tab_px(uint32_t e, const int *P) {
    const int a = P[e & 31], b = P[(e >> 5) & 31], c = P[(e >> 10) & 31];
    return (e >> 15) ? (a + b + 1) >> 1 : (a + 2 * b + c + 2) >> 2;
}

// write one reconstructed sample: pred (or, when the reference would have predicted nothing, what the
// buffer holds — Q15) + residual, clipped.
__device__ __forceinline__ void put_px(uint8_t *p, int have, int pred, int res) {
    if (!have) pred = __ldcg(p);
    *p = (uint8_t)clip255(pred + res);
}

// Reconstruct one intra (or I_PCM) macroblock with one warp.  FAST (progressive pictures): neighbour samples and
// the MB itself live in a shared-memory tile for the whole MB (one L2 round trip in, one out); otherwise every
// sample access goes through the generic MBAFF-aware neighbour derivation.
// PLANES: 1 = luma only, 2 = Cb and Cr only, 3 = everything.  Luma and chroma prediction never read each other's samples, so progressive
// pictures run them as two independent wavefronts (two launches of k_intra side by side): the luma chain loses the chroma work of
// every macroblock, the chroma chain (one step per macroblock, neighbours A, B, D only) is short.
template <bool FAST, int PLANES = 3>
__device__ inline void intra_mb(const PicDev &P, int a, const H264B2MbInfo &I, int lane, IntraWarpSmem &S, const uint16_t *tab4, const uint16_t *tab8) {
    constexpr bool LUMA = (PLANES & 1) != 0, CHROMA = (PLANES & 2) != 0;
    const int field = P.mbaff && (I.flags & H264B2_MBF_FIELD);
    const int ys = field ? 2 : 1;
    int x0, y0;
    mb_origin(P, a, field, x0, y0);
    const int W = P.wmb * 16, H = P.hmb * 16, Wc = W >> 1;
    const int xc0 = x0 >> 1, yc0 = chroma_y0(y0);
    uint8_t *Y = P.dst, *Cb = P.dst + (size_t)W * H, *Cr = Cb + (size_t)Wc * (H >> 1);
    const int cls = I.mb_class;
    if (cls == H264B2_MB_IPCM) {                                       // PB:2449
        if (!mb_coefs_in_bounds(P, a, H264B2_CM_PCM, cls, 0)) return;      // samples beyond the coefficient array: leave the buffer as it is
        const int16_t *pcm = P.coefs + P.coef_off[a];
        if (LUMA) for (int i = lane; i < 256; i += 32) Y[(size_t)(y0 + ys * (i >> 4)) * W + x0 + (i & 15)] = (uint8_t)pcm[i];
        if (CHROMA) for (int i = lane; i < 64; i += 32) {
            Cb[(size_t)(yc0 + ys * (i >> 3)) * Wc + xc0 + (i & 7)] = (uint8_t)pcm[256 + i];
            Cr[(size_t)(yc0 + ys * (i >> 3)) * Wc + xc0 + (i & 7)] = (uint8_t)pcm[320 + i];
        }
        return;
    }
    int avA = 0, avB = 0, avC = 0, avD = 0;
    if (FAST) {
        // availability of the four neighbouring MBs (PB:2887-2947) incl. constrained_intra_pred (PB:1129)
        // geometric neighbours; under MBAFF (frame MBs among frame pairs only, see intra_fast_ok) the address of the
        // MB at (x, y) is 2*((y/2)*w + x) + y%2, and "address <= current" makes C unavailable for bottom MBs (PB:3058-3077)
        const int w = P.wmb, mx = x0 >> 4, my = y0 >> 4;
        auto addr = [&](int x, int y) { return P.mbaff ? 2 * ((y >> 1) * w + x) + (y & 1) : y * w + x; };
        const int nA = mx > 0 ? addr(mx - 1, my) : -1, nB = my > 0 ? addr(mx, my - 1) : -1;
        const int nC = (my > 0 && mx + 1 < w) ? addr(mx + 1, my - 1) : -1, nD = (mx > 0 && my > 0) ? addr(mx - 1, my - 1) : -1;
        avA = nA >= 0 && avail_addr(P, a, nA) && !(P.info[nA].flags & H264B2_MBF_CIP_UNAVAIL);
        avB = nB >= 0 && avail_addr(P, a, nB) && !(P.info[nB].flags & H264B2_MBF_CIP_UNAVAIL);
        avC = nC >= 0 && avail_addr(P, a, nC) && !(P.info[nC].flags & H264B2_MBF_CIP_UNAVAIL);
        avD = nD >= 0 && avail_addr(P, a, nD) && !(P.info[nD].flags & H264B2_MBF_CIP_UNAVAIL);
        // stage the tile: own samples (kept where a block is not predicted, Q15), the row above, the left column
        const uint8_t *Yp = Y + (size_t)y0 * W + x0;
        if (LUMA) {
#pragma unroll
            for (int t = 0; t < 2; t++) { const int wd = lane + 32 * t, r = wd >> 2, j = wd & 3; S.Yt[1 + r][1 + j] = __ldcg((const uint32_t *)(Yp + (size_t)r * W + 4 * j)); }
        }
        if (CHROMA) { const int c = lane >> 4, cl = lane & 15, r = cl >> 1, j = cl & 1;
          S.Ct[c][1 + r][1 + j] = __ldcg((const uint32_t *)((c ? Cr : Cb) + (size_t)(yc0 + r) * Wc + xc0 + 4 * j)); }
        if (y0 > 0) {
            if (lane < 7) { if (LUMA && (lane > 0 || x0 > 0) && (lane < 5 || x0 + 16 < W)) S.Yt[0][lane] = __ldcg((const uint32_t *)(Yp - W - 4 + 4 * lane)); }
            else if (lane >= 8 && lane < 14) { const int l = lane - 8, c = l / 3, j = l % 3;
                if (CHROMA && (j > 0 || x0 > 0)) S.Ct[c][0][j] = __ldcg((const uint32_t *)((c ? Cr : Cb) + (size_t)(yc0 - 1) * Wc + xc0 - 4 + 4 * j)); }
        }
        if (x0 > 0) {
            if (lane < 16) { if (LUMA) ((uint8_t *)S.Yt[1 + lane])[3] = __ldcg(Yp + (size_t)lane * W - 1); }
            else if (CHROMA) { const int l = lane - 16, c = l >> 3, r = l & 7; ((uint8_t *)S.Ct[c][1 + r])[3] = __ldcg((c ? Cr : Cb) + (size_t)(yc0 + r) * Wc + xc0 - 1); }
        }
    }
    auto sample = [&](int xN, int yN, int comp) -> int {
        if (FAST) return tile_sample(S, xN, yN, comp, avA, avB, avC, avD);
        return nbr_sample(P, a, xN, yN, comp);
    };
    auto put = [&](int comp, int x, int y, int have, int pred, int r) {
        if (FAST) {
            uint8_t *px = comp ? &((uint8_t *)S.Ct[comp - 1][1 + y])[4 + x] : &((uint8_t *)S.Yt[1 + y])[4 + x];
            if (!have) pred = *px;
            *px = (uint8_t)clip255(pred + r);
        } else {
            uint8_t *px = comp == 0 ? &Y[(size_t)(y0 + y * ys) * W + x0 + x] : &(comp == 1 ? Cb : Cr)[(size_t)(yc0 + y * ys) * Wc + xc0 + x];
            put_px(px, have, pred, r);
        }
    };
    {   // stage this MB's residual (k_residual wrote all 24 blocks of an intra MB) as a raster tile: luma 16x16, Cb 8x8, Cr 8x8
        const uint32_t *src = (const uint32_t *)(P.res + (size_t)a * RES_MB_STRIDE);
        const int hasres = mb_has_residual(I);
#pragma unroll
        for (int t = 0; t < 6; t++) {
            const int w = lane + 32 * t, e = 2 * w, slot = e >> 4, inner = e & 15;
            if (!(slot < 16 ? LUMA : CHROMA)) continue;
            const uint32_t v = hasres ? src[w] : 0u;
            int dsti;
            if (slot < 16) dsti = ((slot >> 2) * 4 + (inner >> 2)) * 16 + (slot & 3) * 4 + (inner & 3);
            else { const int c = (slot - 16) >> 2, b = (slot - 16) & 3; dsti = 256 + c * 64 + ((b >> 1) * 4 + (inner >> 2)) * 8 + (b & 1) * 4 + (inner & 3); }
            *(uint32_t *)&S.rt.res[dsti] = v;
        }
        __syncwarp();
    }
    const int16_t *res = S.rt.res;

    if (!LUMA) { }
    else if (cls == H264B2_MB_I16x16) {                                      // PB:1847
        const int mode = I.pred16_chroma & 3;
        const int nv = lane < 16 ? sample(lane, -1, 0) : sample(-1, lane - 16, 0);     // lanes 0-15 the row above, 16-31 the left column
        S.nb[1 + lane] = nv;
        if (lane == 0) S.nb[0] = sample(-1, -1, 0);
        const unsigned av = __ballot_sync(0xffffffffu, nv >= 0);
        const int topok = (av & 0xFFFFu) == 0xFFFFu, leftok = (av >> 16) == 0xFFFFu;
        const int sumT = __reduce_add_sync(0xffffffffu, lane < 16 ? nv : 0), sumL = __reduce_add_sync(0xffffffffu, lane < 16 ? 0 : nv);
        __syncwarp();
        const int *top = S.nb + 1, *left = S.nb + 17;
        const int corner = S.nb[0];
        int have = 0, dcv = 128, aa = 0, bb = 0, cc = 0;
        if (mode == 0) have = topok;
        else if (mode == 1) have = leftok;
        else if (mode == 2) {
            have = 1;
            if (topok && leftok) dcv = (sumT + sumL + 16) >> 5;
            else if (leftok) dcv = (sumL + 8) >> 4;
            else if (topok) dcv = (sumT + 8) >> 4;
        } else if (topok && leftok) {                                   // the reference does not test p[-1,-1] here (PB:2014-2018)
            have = 1; int Hh = 0, V = 0;
            for (int i = 0; i < 8; i++) { Hh += (i + 1) * (top[8 + i] - (6 - i >= 0 ? top[6 - i] : corner)); V += (i + 1) * (left[8 + i] - (6 - i >= 0 ? left[6 - i] : corner)); }
            aa = 16 * (left[15] + top[15]); bb = (5 * Hh + 32) >> 6; cc = (5 * V + 32) >> 6;
        }
        for (int i = lane; i < 256; i += 32) {
            const int x = i & 15, y = i >> 4;
            int pred = mode == 0 ? top[x] : mode == 1 ? left[y] : mode == 2 ? dcv : clip255((aa + bb * (x - 7) + cc * (y - 7) + 16) >> 5);
            put(0, x, y, have, pred, res[y * 16 + x]);
        }
    } else if (cls == H264B2_MB_I8x8) {                                 // PB:1423, PB:3651
        const uint64_t modes = P.modes[a];
        for (int b = 0; b < 4; b++) {
            const int mode = (int)((modes >> (4 * b)) & 15);
            const int xO = (b & 1) * 8, yO = (b >> 1) * 8;
            // 25 neighbours, one per lane: lane 0 corner, lanes 1..8 left column, lanes 9..24 the 16 samples above
            int v = -1;
            if (lane < 9) v = sample(xO - 1, yO + lane - 1, 0);
            else if (lane < 25) v = sample(xO + lane - 9, yO - 1, 0);
            unsigned av = __ballot_sync(0xffffffffu, v >= 0);
            const bool sub = !(av & 0x1FE0000u) && (av & 0x10000u);     // no top-right sample, p[7,-1] available: replicate it (PB:1516-1528)
            const int v16 = __shfl_sync(0xffffffffu, v, 16);
            if (sub && lane >= 17 && lane < 25) v = v16;
            if (sub) av |= 0x1FE0000u;
            const int top16 = ((av >> 9) & 0xFFFFu) == 0xFFFFu, left8 = ((av >> 1) & 0xFFu) == 0xFFu, cav = av & 1u;
            {   // reference sample filtering 8.3.2.2.1 (PB:1536-1599): lane i produces filtered sample i from its neighbours' lanes
                const int vm = __shfl_up_sync(0xffffffffu, v, 1), vp = __shfl_down_sync(0xffffffffu, v, 1);
                const int pc = __shfl_sync(0xffffffffu, v, 0), pt0 = __shfl_sync(0xffffffffu, v, 9), pl0 = __shfl_sync(0xffffffffu, v, 1);
                int q = -1;
                if (lane == 0) {
                    if (cav) {
                        const int t0 = (av >> 9) & 1u, l0 = (av >> 1) & 1u;
                        if (!t0 || !l0) q = t0 ? (3*pc + pt0 + 2) >> 2 : l0 ? (3*pc + pl0 + 2) >> 2 : pc;
                        else q = (pt0 + 2*pc + pl0 + 2) >> 2;
                    }
                } else if (lane < 9) {
                    if (left8) q = lane == 1 ? (cav ? (pc + 2*v + vp + 2) >> 2 : (3*v + vp + 2) >> 2) : lane == 8 ? (vm + 3*v + 2) >> 2 : (vm + 2*v + vp + 2) >> 2;
                } else if (lane < 25) {
                    if (top16) q = lane == 9 ? (cav ? (pc + 2*v + vp + 2) >> 2 : (3*v + vp + 2) >> 2) : lane == 24 ? (vm + 3*v + 2) >> 2 : (vm + 2*v + vp + 2) >> 2;
                }
                if (lane < 25) S.qa[lane] = q;
            }
            __syncwarp();
            if (mode >= 3 && mode <= 8) {
                const int have = mode == 8 ? left8 : (mode == 3 || mode == 7) ? top16 : (top16 && left8 && cav);
#pragma unroll
                for (int i = lane; i < 64; i += 32) {
                    const int x = i & 7, y = i >> 3;
                    const int pred = have ? pred_tab_px(tab8[(mode - 3) * 64 + i], S.qa) : 0;
                    put(0, xO + x, yO + y, have, pred, res[(yO + y) * 16 + xO + x]);
                }
            } else {
                for (int i = lane; i < 64; i += 32) {
                    const int x = i & 7, y = i >> 3;
                    int have; const int pred = pred8x8_px(mode, x, y, S.qa, top16, top16, left8, cav, have);
                    put(0, xO + x, yO + y, have, pred, res[(yO + y) * 16 + xO + x]);
                }
            }
            __syncwarp();
        }
    } else {                                                            // Intra_4x4: PB:1062, PB:3401
        const uint64_t modes = P.modes[a];
        for (int b = 0; b < 16; b++) {
            const int mode = (int)((modes >> (4 * b)) & 15);
            const int xO = blk_x(b), yO = blk_y(b);
            if (lane < 5) S.nb[lane] = sample(xO - 1, yO + lane - 1, 0);
            else if (lane < 13) { const int x = lane - 5; S.nb[lane] = (x > 3 && (b == 3 || b == 11)) ? -1 : sample(xO + x, yO - 1, 0); }
            __syncwarp();
            if (lane == 0 && S.nb[9] < 0 && S.nb[10] < 0 && S.nb[11] < 0 && S.nb[12] < 0 && S.nb[8] >= 0) { S.nb[9] = S.nb[10] = S.nb[11] = S.nb[12] = S.nb[8]; }
            __syncwarp();
            if (lane < 16) {
                const int x = lane & 3, y = lane >> 2;
                int have, pred;
                if (mode >= 3 && mode <= 8) {
                    const int topok = S.nb[5] >= 0 && S.nb[6] >= 0 && S.nb[7] >= 0 && S.nb[8] >= 0, trok = S.nb[9] >= 0 && S.nb[10] >= 0 && S.nb[11] >= 0 && S.nb[12] >= 0;
                    const int leftok = S.nb[1] >= 0 && S.nb[2] >= 0 && S.nb[3] >= 0 && S.nb[4] >= 0, cornok = S.nb[0] >= 0;
                    have = mode == 8 ? leftok : (mode == 3 || mode == 7) ? (topok && trok) : (topok && leftok && cornok);
                    pred = have ? pred_tab_px(tab4[(mode - 3) * 16 + lane], S.nb) : 0;
                } else pred = pred4x4_px(mode, x, y, S.nb, have);
                put(0, xO + x, yO + y, have, pred, res[(yO + y) * 16 + xO + x]);
            }
            __syncwarp();
        }
    }

    // chroma (PB:2076): the two components are independent, so lanes 0-15 predict Cb while lanes 16-31 predict Cr
    if (CHROMA) {
        const int cmode = (I.pred16_chroma >> 2) & 3;
        const int comp = 1 + (lane >> 4), hl = lane & 15;
        int *nbc = S.nb + 17 * (lane >> 4);              // [0] corner, [1..8] top, [9..16] left
        __syncwarp();
        const int nv = hl < 8 ? sample(hl, -1, comp) : sample(-1, hl - 8, comp);
        nbc[1 + hl] = nv;                                    // [1..8] top, [9..16] left
        if (hl == 0) nbc[0] = sample(-1, -1, comp);
        const unsigned av = (__ballot_sync(0xffffffffu, nv >= 0) >> (lane & 16)) & 0xFFFFu;
        const int topok = (av & 0xFFu) == 0xFFu, leftok = (av >> 8) == 0xFFu;
        __syncwarp();
        const int *top = nbc + 1, *left = nbc + 9;
        const int corner = nbc[0];
        const int16_t *cres = res + 256 + (comp - 1) * 64;
        int have = 0, aa = 0, bb = 0, cc = 0;
        if (cmode == 0) have = 1;
        else if (cmode == 1) have = leftok;
        else if (cmode == 2) have = topok;
        else if (topok && leftok && corner >= 0) {
            have = 1; int Hh = 0, V = 0;
            for (int i = 0; i < 4; i++) { Hh += (i + 1) * (top[4 + i] - (2 - i >= 0 ? top[2 - i] : corner)); V += (i + 1) * (left[4 + i] - (2 - i >= 0 ? left[2 - i] : corner)); }
            aa = 16 * (left[7] + top[7]); bb = (34 * Hh + 32) >> 6; cc = (34 * V + 32) >> 6;
        }
        for (int i = hl; i < 64; i += 16) {
            const int x = i & 7, y = i >> 3;
            int pred;
            if (cmode == 0) {
                const int xO = x & 4, yO = y & 4;
                // `> 0`: a neighbouring sample equal to 0 counts as unavailable (Q3, PB:2205-2250)
                const int t = top[xO] > 0 && top[xO+1] > 0 && top[xO+2] > 0 && top[xO+3] > 0;
                const int l = left[yO] > 0 && left[yO+1] > 0 && left[yO+2] > 0 && left[yO+3] > 0;
                const int st = top[xO] + top[xO+1] + top[xO+2] + top[xO+3], sl = left[yO] + left[yO+1] + left[yO+2] + left[yO+3];
                if ((xO == 0 && yO == 0) || (xO > 0 && yO > 0)) pred = (t && l) ? (st + sl + 4) >> 3 : l ? (sl + 2) >> 2 : t ? (st + 2) >> 2 : 128;
                else if (xO > 0) pred = t ? (st + 2) >> 2 : l ? (sl + 2) >> 2 : 128;
                else pred = l ? (sl + 2) >> 2 : t ? (st + 2) >> 2 : 128;
            } else if (cmode == 1) pred = left[y];
            else if (cmode == 2) pred = top[x];
            else pred = clip255((aa + bb * (x - 3) + cc * (y - 3) + 16) >> 5);
            put(comp, x, y, have, pred, cres[y * 8 + x]);
        }
    }
    __syncwarp();
    if (FAST) {
        uint8_t *Yp = Y + (size_t)y0 * W + x0;
        if (LUMA) {
#pragma unroll
            for (int t = 0; t < 2; t++) { const int wd = lane + 32 * t, r = wd >> 2, j = wd & 3; *(uint32_t *)(Yp + (size_t)r * W + 4 * j) = S.Yt[1 + r][1 + j]; }
        }
        if (CHROMA) {
            const int c = lane >> 4, cl = lane & 15, r = cl >> 1, j = cl & 1;
            *(uint32_t *)((c ? Cr : Cb) + (size_t)(yc0 + r) * Wc + xc0 + 4 * j) = S.Ct[c][1 + r][1 + j];
        }
    }
}

// May macroblock `a` of an MBAFF picture take the staged (geometric) path?  Yes when it is a frame MB and every pair
// it can take neighbour samples from (left, above-left, above, above-right) is a frame pair too: then 6.4.12.2
// degenerates to the geometric neighbours of a progressive picture.
__device__ __forceinline__ bool intra_fast_ok(const PicDev &P, int a) {
    if (!P.mbaff) return true;
    if (P.info[a].flags & H264B2_MBF_FIELD) return false;
    const int w = P.wmb, pr = a >> 1, px = pr % w, py = pr / w;
    bool ok = true;
    if (px > 0) ok &= !(P.info[2 * (pr - 1)].flags & H264B2_MBF_FIELD);
    if (py > 0) {
        ok &= !(P.info[2 * (pr - w)].flags & H264B2_MBF_FIELD);
        if (px > 0) ok &= !(P.info[2 * (pr - w - 1)].flags & H264B2_MBF_FIELD);
        if (px + 1 < w) ok &= !(P.info[2 * (pr - w + 1)].flags & H264B2_MBF_FIELD);
    }
    return ok;
}

// Wavefront driver (see wavefront.cuh): one CTA per band of WF_ROWS MB rows, one warp per row.
// PLANES (progressive pictures only): 3 = one wavefront for luma and chroma, 1 / 2 = the luma / chroma wavefront of the split schedule
// (own progress counters: progress[0][row] luma or both, progress[2][row] chroma).
template <bool GENERIC, int PLANES = 3>
__global__ void __launch_bounds__(WF_THREADS, 1024 / WF_THREADS) k_intra(const PicDev *pics, int npics, int bands, int *ticket) {
    __shared__ IntraWarpSmem sm[WF_ROWS];
    __shared__ int s_prog[WF_ROWS];
    __shared__ uint64_t s_bar[WF_ROWS];
    __shared__ uint32_t s_mask[WF_ROWS][2][8];
    __shared__ int s_ticket;
    __shared__ uint16_t s_tab4[6 * 16], s_tab8[6 * 64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 6 * 64; i += blockDim.x) { s_tab8[i] = (&g_pred8_tab[0][0])[i]; if (i < 6 * 16) s_tab4[i] = (&g_pred4_tab[0][0])[i]; }
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    if (threadIdx.x < WF_ROWS) { s_prog[threadIdx.x] = 0; mbar_init(&s_bar[threadIdx.x], 1); }
    __syncthreads();
    const int t = s_ticket;
    if (t >= npics * bands) return;
    const PicDev &P = pics[t % npics];
    const int row = (t / npics) * WF_ROWS + warp;
    const int per = P.mbaff ? 2 : 1;
    const int rows = P.hmb / per, wmb = P.wmb;
    if (row >= rows) return;
    RowSync rs = rs_init(s_prog, s_bar, warp, row, rows, P.progress + (PLANES == 2 ? 2 * P.hmb : 0), wmb);      // progress[0][row], chroma wavefront: progress[2][row]
    // Intra masks of this row and of the row above.  An intra MB only has to wait for the row above if one of its
    // neighbours B, C, D there is itself intra: inter neighbours were completed by k_inter before this kernel
    // started.  In P/B pictures, where intra MBs are scattered, this removes the false chains a plain
    // "row above >= x+2" rule would build through unrelated macroblocks.
    uint32_t *mine = s_mask[warp][0], *above = s_mask[warp][1];
    const bool masks_ok = wmb <= 256;
    for (int g = 0; g < 8 && masks_ok; g++) {
        const int xl = g * 32 + lane;
        int im = 0, ia = 0;
        if (xl < wmb) for (int s = 0; s < per; s++) {
            const int c = P.info[(row * wmb + xl) * per + s].mb_class; im |= (c >= H264B2_MB_I4x4 && c <= H264B2_MB_IPCM);
            if (row > 0) { const int ca = P.info[((row - 1) * wmb + xl) * per + s].mb_class; ia |= (ca >= H264B2_MB_I4x4 && ca <= H264B2_MB_IPCM); }
        }
        const unsigned mm = __ballot_sync(0xffffffffu, im), ma = __ballot_sync(0xffffffffu, ia);
        if (lane == 0) { mine[g] = mm; above[g] = ma; }
    }
    __syncwarp();
    auto above_intra = [&](int x) -> bool { return x >= 0 && x < wmb && ((above[x >> 5] >> (x & 31)) & 1u); };
    // Side information of the NEXT intra macroblock of this row (residual tile, its own and its neighbours' info records, pred
    // modes) is pulled into L1 while the current one is reconstructed: none of it is written during this launch, and the
    // per-macroblock chain (wait -> info -> neighbour info -> residual -> predict) is what bounds this kernel.
    auto prefetch_mb = [&](int x) {
        if (GENERIC || x < 0 || x >= wmb) return;
        const int a = row * wmb + x;
        const void *p = nullptr;
        if (lane < 6) p = (const uint8_t *)(P.res + (size_t)a * RES_MB_STRIDE) + lane * 128;
        else if (lane == 6) p = &P.info[a];
        else if (lane == 7 && x > 0) p = &P.info[a - 1];
        else if (lane == 8 && row > 0) p = &P.info[a - wmb + (x + 1 < wmb ? 1 : 0)];
        else if (lane == 9 && row > 0) p = &P.info[a - wmb - (x > 0 ? 1 : 0)];
        else if (lane == 10) p = &P.modes[a];
        if (p) asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
    };
    auto next_intra = [&](int x) -> int {          // first intra macroblock of this row after column x, or wmb
        if (!masks_ok) return wmb;
        for (int i = x + 1; i < wmb; ) { const uint32_t w = mine[i >> 5] >> (i & 31); if (w) return i + __ffs(w) - 1; i = (i | 31) + 1; }
        return wmb;
    };
    prefetch_mb(next_intra(-1));
    for (int xb = 0; xb < wmb; xb += 32) {
        unsigned mask;
        if (masks_ok) mask = mine[xb >> 5];
        else {
            const int xl = xb + lane;
            int intra_here = 0;
            if (xl < wmb) for (int s = 0; s < per; s++) { const int c = P.info[(row * wmb + xl) * per + s].mb_class; intra_here |= (c >= H264B2_MB_I4x4 && c <= H264B2_MB_IPCM); }
            mask = __ballot_sync(0xffffffffu, intra_here);
        }
        while (mask) {
            const int x = xb + __ffs(mask) - 1;
            mask &= mask - 1;
            int need = min(x + 2, wmb);
            if (masks_ok) need = (PLANES != 2 && above_intra(x + 1)) ? x + 2 : above_intra(x) ? x + 1 : above_intra(x - 1) ? x : 0;      // chroma prediction has no top-right neighbour
            else if (PLANES == 2) need = min(x + 1, wmb);
            prefetch_mb(next_intra(x));
            rs_wait(rs, need, x, lane);
            for (int s = 0; s < per; s++) {
                const int a = (row * wmb + x) * per + s;
                const H264B2MbInfo I = P.info[a];
                if (I.mb_class >= H264B2_MB_I4x4 && I.mb_class <= H264B2_MB_IPCM) {
                    if (!GENERIC || intra_fast_ok(P, a)) intra_mb<true, PLANES>(P, a, I, lane, sm[warp], s_tab4, s_tab8); else intra_mb<false, 3>(P, a, I, lane, sm[warp], s_tab4, s_tab8);
                }
            }
            rs_publish(rs, x + 1, lane);
        }
    }
    rs_publish(rs, wmb, lane);
}
