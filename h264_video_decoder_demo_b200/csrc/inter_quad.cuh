// inter_quad.cuh — k_inter: inter prediction + inter residual add, one WARP per macroblock, 8 lanes per 8x8 quadrant.
//
// Reference: Inter_prediction_process IP:412-667 (partition walk), fractional sample interpolation IP:2228-2328,
// luma 6-tap IP:2344-2480, chroma bilinear IP:2485-2522, weighted prediction IP:2526-2829, residual add IP:22-407.
//
// Design (B200).  The reference interpolates sample by sample; the first version of this kernel ran one thread per
// 4x4 block (inter.cuh), which loads a 9x9 window per 16 samples and makes the lanes of a warp execute the union of
// every fractional-position class they hold.  Motion in real streams is far coarser than 4x4: in the bundled 1080p
// streams every 8x8 quadrant of every inter macroblock carries ONE vector per list (81 % of the macroblocks carry one
// vector for all 256 samples).  So the unit of work here is the quadrant:
//   * lane (q, r) = quadrant q (0..3), row r (0..7).  The 8 lanes of a quadrant share reference picture, vector and
//     fractional position: no divergence inside the group, and groups only diverge when the quadrants really differ.
//   * the 13x13 luma window of the quadrant is read ONCE: lane r loads window rows r and r+8 as 4 aligned 32-bit words
//     each (funnel-shifted to byte alignment), keeps them in shared memory, and — when the position has a horizontal
//     component — also leaves the unclipped horizontal 6-tap sums of its rows there (DP4A on packed bytes);
//   * after one group-level __syncwarp, lane r produces output row r: horizontal half samples from the stored sums,
//     vertical half samples by a packed 2x16-bit column filter over 6 window rows, the centre position j from 6 rows
//     of horizontal sums in 32-bit; quarter positions are byte-wise rounded averages of two of those;
//   * chroma: lane (q, r) produces row r&3 of plane r>>2 of the quadrant's 4x4 chroma block with 4 DP4A;
//   * stores: 8 bytes of luma per lane (quadrant rows are 16-byte contiguous per macroblock row), 4 bytes of chroma.
// Macroblocks with sub-8x8 motion take the per-4x4-block routine of inter.cuh on lanes 0..15 (bit-exact, rarely used).
#pragma once
#include "inter.cuh"

struct alignas(16) InterWarpSmem {
    uint32_t raw[4][13][4];   // per quadrant: window rows, byte b <-> x = xI - 2 + b (13 of 16 bytes used)
    int      hs[4][13][8];    // unclipped horizontal 6-tap sums of the window rows, output x = 0..7
};

__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack4_clip(int a, int b, int c, int d, int rnd, int sh) {
    return pack4(clip255((a + rnd) >> sh), clip255((b + rnd) >> sh), clip255((c + rnd) >> sh), clip255((d + rnd) >> sh));
}

// Output row r of a quadrant from its staged window rows (raw) and their unclipped horizontal sums (hs): the half-sample
// intermediates b/s, h/m, j and the quarter-sample averages of IP:2415-2477.
__device__ __forceinline__ void luma_quad_combine(const uint32_t (*raw)[4], const int (*hs)[8], int xF, int yF, int r, uint32_t out[2]) {
    const bool needJ = (xF == 2 && yF != 0) || (yF == 2 && xF != 0);
    const bool needH = xF != 0 && yF != 2;
    const bool needV = yF != 0 && xF != 2;
    const int ry = 2 + (yF == 3);
    uint32_t G[2] = {0, 0}, Hh[2] = {0, 0}, Vh[2] = {0, 0}, J[2] = {0, 0};
    if (xF == 0 || yF == 0) {                               // full-sample row next to the fractional position
        const uint4 A = *(const uint4 *)raw[r + 2 + (xF == 0 && yF == 3)];
        const uint32_t sel = (yF == 0 && xF == 3) ? 0x6543u : 0x5432u;
        G[0] = __byte_perm(A.x, A.y, sel); G[1] = __byte_perm(A.y, A.z, sel);
    }
    if (needH) {                                            // b (row y) or s (row y+1)
        const int4 ha = *(const int4 *)&hs[r + ry][0], hb = *(const int4 *)&hs[r + ry][4];
        Hh[0] = pack4_clip(ha.x, ha.y, ha.z, ha.w, 16, 5); Hh[1] = pack4_clip(hb.x, hb.y, hb.z, hb.w, 16, 5);
    }
    if (needV) {                                            // h (column x) or m (column x+1): 6 rows, packed 2 x 16 bit
        const uint32_t sel = xF == 3 ? 0x6543u : 0x5432u;
        uint32_t acc[4] = {0x0A100A10u, 0x0A100A10u, 0x0A100A10u, 0x0A100A10u};    // bias 2560 keeps every half non-negative, +16 rounding
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const uint4 A = *(const uint4 *)raw[r + k];
            const uint32_t g0 = __byte_perm(A.x, A.y, sel), g1 = __byte_perm(A.y, A.z, sel);
            const uint32_t e[4] = { __byte_perm(g0, 0, 0x4140), __byte_perm(g0, 0, 0x4342), __byte_perm(g1, 0, 0x4140), __byte_perm(g1, 0, 0x4342) };
            const int c = (k == 0 || k == 5) ? 1 : (k == 1 || k == 4) ? -5 : 20;
#pragma unroll
            for (int i = 0; i < 4; i++) acc[i] += (uint32_t)c * e[i];
        }
        uint32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t t = (acc[i] >> 5) & 0x07FF07FFu;                        // ((sum + 16) >> 5) + 80 per half
            v[i] = __vminu2(__vmaxu2(t, 0x00500050u), 0x014F014Fu) - 0x00500050u;   // Clip1
        }
        Vh[0] = __byte_perm(v[0], v[1], 0x6420); Vh[1] = __byte_perm(v[2], v[3], 0x6420);
    }
    if (needJ) {                                            // j: 6-tap down the unclipped horizontal sums
        int t[8];
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = 512;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int4 ha = *(const int4 *)&hs[r + k][0], hb = *(const int4 *)&hs[r + k][4];
            const int c = (k == 0 || k == 5) ? 1 : (k == 1 || k == 4) ? -5 : 20;
            t[0] += c * ha.x; t[1] += c * ha.y; t[2] += c * ha.z; t[3] += c * ha.w;
            t[4] += c * hb.x; t[5] += c * hb.y; t[6] += c * hb.z; t[7] += c * hb.w;
        }
        J[0] = pack4_clip(t[0], t[1], t[2], t[3], 0, 10); J[1] = pack4_clip(t[4], t[5], t[6], t[7], 0, 10);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        uint32_t o;
        if (xF == 0 && yF == 0) o = G[i];
        else if (xF == 0) o = yF == 2 ? Vh[i] : avg4(G[i], Vh[i]);            // d, h, n
        else if (yF == 0) o = xF == 2 ? Hh[i] : avg4(G[i], Hh[i]);            // a, b, c
        else if (xF == 2 && yF == 2) o = J[i];
        else if (xF == 2) o = avg4(J[i], Hh[i]);                              // f, q
        else if (yF == 2) o = avg4(J[i], Vh[i]);                              // i, k
        else o = avg4(Hh[i], Vh[i]);                                          // e, g, p, r
        out[i] = o;
    }
}

// Predicted luma row r (8 samples, two packed words) of one quadrant for one list.  (xI, yI): integer position of the
// quadrant's top-left sample in the reference view.  Called by the 8 lanes of a quadrant together (gmask).
__device__ __forceinline__ void luma_quad_pred(const uint8_t *base, int stride, int wclamp, int hclamp, int wfast,
                                               int xI, int yI, int xF, int yF, int r, unsigned gmask,
                                               uint32_t (*raw)[4], int (*hs)[8], uint32_t out[2]) {
    const bool fast = xI >= 2 && xI + 14 <= wfast && yI >= 2 && yI + 10 < hclamp;
    const bool needJ = (xF == 2 && yF != 0) || (yF == 2 && xF != 0);
    const bool needH = xF != 0 && yF != 2;
    const bool needV = yF != 0 && xF != 2;
    const int ry = 2 + (yF == 3);
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int w = r + 8 * t;
        if (w < 13 && (yF != 0 || (w >= 2 && w < 10))) {
            uint32_t A0, A1, A2, A3;
            if (fast) {
                const uint8_t *p = base + (size_t)(yI - 2 + w) * stride + (xI - 2);
                const int o = (int)((uintptr_t)p & 3);
                const uint32_t *pw = (const uint32_t *)(p - o);
                const int sh = o * 8;
                const uint32_t w0 = __ldg(pw), w1 = __ldg(pw + 1), w2 = __ldg(pw + 2), w3 = __ldg(pw + 3);
                A0 = __funnelshift_r(w0, w1, sh); A1 = __funnelshift_r(w1, w2, sh); A2 = __funnelshift_r(w2, w3, sh); A3 = w3 >> sh;
            } else {
                const uint8_t *row = base + (size_t)clip3i(0, hclamp - 1, yI - 2 + w) * stride;
                uint32_t b[13];
#pragma unroll
                for (int i = 0; i < 13; i++) b[i] = __ldg(row + clip3i(0, wclamp - 1, xI - 2 + i));
                A0 = pack4(b[0], b[1], b[2], b[3]); A1 = pack4(b[4], b[5], b[6], b[7]); A2 = pack4(b[8], b[9], b[10], b[11]); A3 = b[12];
            }
            *(uint4 *)raw[w] = make_uint4(A0, A1, A2, A3);
            if (xF != 0 && (needJ || (w >= ry && w < ry + 8))) {
                int4 ha, hb;
                tap6x4(A0, A1, A2, ha.x, ha.y, ha.z, ha.w);
                tap6x4(A1, A2, A3, hb.x, hb.y, hb.z, hb.w);
                *(int4 *)&hs[w][0] = ha; *(int4 *)&hs[w][4] = hb;
            }
        }
    }
    __syncwarp(gmask);
    luma_quad_combine(raw, hs, xF, yF, r, out);
    __syncwarp(gmask);                                       // the tile is reused by the next list
}

// Row cy (4 samples) of the quadrant's 4x4 chroma block of one plane for one list (IP:2485-2522)
__device__ __forceinline__ void chroma_quad_pred(const uint8_t *base, int stride, int wclamp, int hclamp, int wfast,
                                                 int xC, int yC, int xF, int yF, int p[4]) {
    uint32_t s[2], e[2];                                     // bytes 0..3 and byte 4 of rows yC, yC + 1
    if (xC >= 0 && xC + 8 <= wfast && yC >= 0 && yC + 1 < hclamp) {
        const uint8_t *q = base + (size_t)yC * stride + xC;
        const int o = (int)((uintptr_t)q & 3);
        const uint32_t *pw = (const uint32_t *)(q - o);
        const int sw = stride >> 2, sh = o * 8;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint32_t w0 = __ldg(pw + k * sw), w1 = __ldg(pw + k * sw + 1);
            s[k] = __funnelshift_r(w0, w1, sh); e[k] = (w1 >> sh) & 0xffu;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint8_t *row = base + (size_t)clip3i(0, hclamp - 1, yC + k) * stride;
            uint32_t b[5];
#pragma unroll
            for (int i = 0; i < 5; i++) b[i] = __ldg(row + clip3i(0, wclamp - 1, xC + i));
            s[k] = pack4(b[0], b[1], b[2], b[3]); e[k] = b[4];
        }
    }
    const uint32_t cf = (uint32_t)((8 - xF) * (8 - yF)) | ((uint32_t)(xF * (8 - yF)) << 8) | ((uint32_t)((8 - xF) * yF) << 16) | ((uint32_t)(xF * yF) << 24);
    const uint32_t h0 = (s[0] >> 24) | (e[0] << 8), h1 = (s[1] >> 24) | (e[1] << 8);
    p[0] = dp4a_uu(__byte_perm(s[0], s[1], 0x5410), cf, 32) >> 6;
    p[1] = dp4a_uu(__byte_perm(s[0], s[1], 0x6521), cf, 32) >> 6;
    p[2] = dp4a_uu(__byte_perm(s[0], s[1], 0x7632), cf, 32) >> 6;
    p[3] = dp4a_uu(__byte_perm(h0, h1, 0x5410), cf, 32) >> 6;
}

// Weighted combination of the lists, residual add and store of this lane's luma row and chroma row (IP:2526-2829, IP:22-407):
// the tail shared by the clamped routine below and by the TMA-staged kernel (inter_tma.cuh).
__device__ __forceinline__ void inter_quad_store(const PicDev &P, int a, uint32_t m, int t8, int x0, int y0, int ys, int q, int r,
                                                 const uint32_t (&pl)[2][2], const int (&pc)[2][4], int have0, int have1, int wt_idx) {
    const int qx = (q & 1) * 8, qy = (q >> 1) * 8, cpl = r >> 2, cy = r & 3;
    const int W = P.wmb * 16, H = P.hmb * 16, Wc = W >> 1;
    const int none = !have0 && !have1;        // the reference predicts nothing: the residual lands on what the buffer holds
    const H264B2Weight w = P.weights[min(wt_idx, P.n_weights - 1)];
    const int16_t *res = P.res + (size_t)a * RES_MB_STRIDE;
    // ---- luma row qy + r, columns qx .. qx+7
    {
        uint8_t *Y = P.dst + (size_t)(y0 + (qy + r) * ys) * W + x0 + qx;
        uint32_t o[2];
        // bi-prediction with w0 = w1 = 2^logWD and a zero offset term IS the default average (IP:2699 with those values)
        const bool plain = !w.mode || (have0 && have1 && w.w0[0] == w.w1[0] && w.w0[0] == (1 << w.logwd[0]) && ((w.o0[0] + w.o1[0] + 1) >> 1) == 0);
        if (none) { const uint2 t = *(const uint2 *)Y; o[0] = t.x; o[1] = t.y; }
        else if (plain) {
#pragma unroll
            for (int i = 0; i < 2; i++) o[i] = (have0 && have1) ? avg4(pl[0][i], pl[1][i]) : have0 ? pl[0][i] : pl[1][i];
        } else {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                int v[4];
#pragma unroll
                for (int x = 0; x < 4; x++) v[x] = weigh(w, 0, have0, have1, (pl[0][i] >> (8 * x)) & 0xff, (pl[1][i] >> (8 * x)) & 0xff);
                o[i] = pack4(v[0], v[1], v[2], v[3]);
            }
        }
        const int slot = ((qy + r) >> 2) * 4 + (qx >> 2);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (luma_slot_coded(m, H264B2_MB_INTER, t8, slot + i)) {
                const uint2 e = *(const uint2 *)(res + (slot + i) * 16 + (r & 3) * 4);
                const int r0 = (int16_t)(e.x & 0xffff), r1 = (int16_t)(e.x >> 16), r2 = (int16_t)(e.y & 0xffff), r3 = (int16_t)(e.y >> 16);
                o[i] = pack4(clip255((int)(o[i] & 0xff) + r0), clip255((int)((o[i] >> 8) & 0xff) + r1), clip255((int)((o[i] >> 16) & 0xff) + r2), clip255((int)(o[i] >> 24) + r3));
            }
        }
        *(uint2 *)Y = make_uint2(o[0], o[1]);
    }
    // ---- chroma plane cpl, row qy/2 + cy, columns qx/2 .. qx/2+3
    {
        uint8_t *C = P.dst + (size_t)W * H + (cpl ? (size_t)Wc * (H >> 1) : 0) + (size_t)(chroma_y0(y0) + (qy / 2 + cy) * ys) * Wc + (x0 >> 1) + qx / 2;
        int v[4];
        if (none) { const uint32_t t = *(const uint32_t *)C; v[0] = t & 0xff; v[1] = (t >> 8) & 0xff; v[2] = (t >> 16) & 0xff; v[3] = t >> 24; }
        else {
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = cpl ? weigh(w, 2, have0, have1, pc[0][i], pc[1][i]) : weigh(w, 1, have0, have1, pc[0][i], pc[1][i]);
        }
        if (chroma_blk_coded(m, cpl, q)) {
            const uint2 e = *(const uint2 *)(res + (16 + cpl * 4 + q) * 16 + cy * 4);
            v[0] = clip255(v[0] + (int16_t)(e.x & 0xffff)); v[1] = clip255(v[1] + (int16_t)(e.x >> 16));
            v[2] = clip255(v[2] + (int16_t)(e.y & 0xffff)); v[3] = clip255(v[3] + (int16_t)(e.y >> 16));
        }
        *(uint32_t *)C = pack4(v[0], v[1], v[2], v[3]);
    }
}

// Does a progressive frame macroblock with ONE vector per list (mv0 / mv1, reference codes code0 / code1; -1 = list unused) take the
// TMA-staged path of inter_tma.cuh?  Frame views only, and every window (21 x 21 luma, 9 x 9 chroma) inside the picture: TMA fills
// what lies outside with zeros, the reference clamps the coordinates (IP:2363).  Both kernels evaluate this same function.
__device__ __forceinline__ bool inter_staged_ok(const PicDev &P, int mbx, int mby, int code0, int code1, uint32_t mv0, uint32_t mv1) {
    if (P.mbaff || P.generic || (code0 < 0 && code1 < 0)) return false;
    if ((code0 >= 0 && (code0 & 3)) || (code1 >= 0 && (code1 & 3))) return false;
    const int W = P.wmb * 16, H = P.hmb * 16, x0 = mbx * 16, y0 = mby * 16;
    bool inside = true;
#pragma unroll
    for (int l = 0; l < 2; l++) {
        if ((l ? code1 : code0) < 0) continue;
        const uint32_t mv = l ? mv1 : mv0;
        const int mvx = (int16_t)(mv & 0xffffu), mvy = (int16_t)(mv >> 16);
        const int xI = x0 + (mvx >> 2), yI = y0 + (mvy >> 2), xC = (x0 >> 1) + (mvx >> 3), yC = (y0 >> 1) + (mvy >> 3);
        inside = inside && xI >= 2 && xI + 18 < W && yI >= 2 && yI + 18 < H && xC >= 0 && xC + 8 < (W >> 1) && yC >= 0 && yC + 8 < (H >> 1);
    }
    return inside;
}

// One inter macroblock by the whole warp with clamped window loads from global memory (any vector, any view, any position).
__device__ __forceinline__ void inter_mb_ldg(const PicDev &P, int mbx, int mby, int lane, uint32_t (*raw)[13][4], int (*hs)[13][8], int skip_staged) {
    const int a = P.mbaff ? 2 * ((mby >> 1) * P.wmb + mbx) + (mby & 1) : mby * P.wmb + mbx;
    const H264B2MbInfo I = P.info[a];
    if (I.mb_class != H264B2_MB_INTER) return;
    const H264B2MbMotion &M = P.motion[a];
    const int q = lane >> 3, r = lane & 7, qx = (q & 1) * 8, qy = (q >> 1) * 8;
    const int b0 = (q >> 1) * 8 + (q & 1) * 2;              // raster slot of the quadrant's first 4x4 block
    const int code0 = M.ref_surf[0][q], code1 = M.ref_surf[1][q];
    // one vector per list in this quadrant?  (mv[l][slot] = two int16 = one 32-bit word; slots b0, b0+1, b0+4, b0+5)
    const uint32_t *mvw = (const uint32_t *)&M.mv[0][0][0];
    uint32_t mv01[2];
    bool uni = true;
#pragma unroll
    for (int l = 0; l < 2; l++) {
        const uint2 t0 = *(const uint2 *)(mvw + l * 16 + b0), t1 = *(const uint2 *)(mvw + l * 16 + b0 + 4);
        mv01[l] = t0.x;
        if ((l ? code1 : code0) >= 0) uni = uni && t0.y == t0.x && t1.x == t0.x && t1.y == t0.x;
    }
    if (!__all_sync(0xffffffffu, uni)) {
        if (lane < 16) inter_block_generic(P, a, P.info[a], lane);
        return;
    }
    if (skip_staged) {
        // macroblocks that k_inter_tma reconstructs (one vector, one reference pair, one weight entry for the whole macroblock, windows
        // inside the picture) are not touched here
        const uint32_t key0 = code0 >= 0 ? mv01[0] : 0u, key1 = code1 >= 0 ? mv01[1] : 0u, wti = M.wt_idx[q];
        // every shuffle is executed by all lanes (no short-circuit in front of a warp-collective)
        const int c0 = __shfl_sync(0xffffffffu, code0, 0), c1 = __shfl_sync(0xffffffffu, code1, 0);
        const uint32_t k0 = __shfl_sync(0xffffffffu, key0, 0), k1 = __shfl_sync(0xffffffffu, key1, 0), w0 = __shfl_sync(0xffffffffu, wti, 0);
        const uint32_t m0 = __shfl_sync(0xffffffffu, mv01[0], 0), m1 = __shfl_sync(0xffffffffu, mv01[1], 0);
        const bool same = (code0 == c0) & (code1 == c1) & (key0 == k0) & (key1 == k1) & (wti == w0);
        const bool all_same = __all_sync(0xffffffffu, same);
        if (all_same && inter_staged_ok(P, mbx, mby, c0, c1, m0, m1)) return;
    }
    const int field = P.mbaff && (I.flags & H264B2_MBF_FIELD);
    const int ys = field ? 2 : 1;
    const int x0 = mbx * 16, y0 = !P.mbaff ? mby * 16 : field ? (mby >> 1) * 32 + (mby & 1) : mby * 16;      // mb_origin()
    const int yA = field ? y0 / 2 : y0;                     // IP:577-580
    const int W = P.wmb * 16, Wc = W >> 1;
    const unsigned gmask = 0xFFu << (q * 8);
    const int cpl = r >> 2, cy = r & 3;                     // this lane's chroma plane and chroma row inside the quadrant's block

    uint32_t pl[2][2] = {{0, 0}, {0, 0}};
    int pc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    const int have0 = code0 >= 0, have1 = code1 >= 0;
#pragma unroll 1
    for (int l = 0; l < 2; l++) {
        const int code = l ? code1 : code0;
        if (code < 0) continue;
        RefViewDev rv;
        ref_view(P, code, rv);
        const uint32_t mvp = l ? mv01[1] : mv01[0];
        const int mvx = (int16_t)(mvp & 0xffff), mvy = (int16_t)(mvp >> 16);
        int mvcy = mvy;
        if (field) { if (rv.view == 1 && (a & 1)) mvcy += 2; else if (rv.view == 2 && !(a & 1)) mvcy -= 2; }   // IP:2019-2043
        uint32_t tl[2];
        int tc[4];
        luma_quad_pred(rv.base[0], rv.stride[0], rv.wclamp[0], rv.hclamp[0], W, x0 + qx + (mvx >> 2), yA + qy + (mvy >> 2), mvx & 3, mvy & 3,
                       r, gmask, raw[q], hs[q], tl);
        const int xC = (x0 + qx) / 2 + (mvx >> 3), yC = (yA + qy) / 2 + (mvcy >> 3) + cy;
        chroma_quad_pred(cpl ? rv.base[2] : rv.base[1], rv.stride[1], rv.wclamp[1], rv.hclamp[1], Wc, xC, yC, mvx & 7, mvcy & 7, tc);     // Cb and Cr share their geometry
        if (l == 0) { pl[0][0] = tl[0]; pl[0][1] = tl[1]; pc[0][0] = tc[0]; pc[0][1] = tc[1]; pc[0][2] = tc[2]; pc[0][3] = tc[3]; }
        else        { pl[1][0] = tl[0]; pl[1][1] = tl[1]; pc[1][0] = tc[0]; pc[1][1] = tc[1]; pc[1][2] = tc[2]; pc[1][3] = tc[3]; }
    }
    inter_quad_store(P, a, I.coef_mask, (I.flags & H264B2_MBF_T8x8) != 0, x0, y0, ys, q, r, pl, pc, have0, have1, M.wt_idx[q]);
}

// grid: (ceil(wmb / 4), hmb, n_pics); block: 128 threads = 4 warps = 4 horizontally consecutive macroblocks (no division
// to find the macroblock: under MBAFF row 2k / 2k+1 are the top / bottom macroblocks of pair row k).
#ifndef INTER_MIN_BLOCKS
#define INTER_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(128, INTER_MIN_BLOCKS) k_inter(const PicDev *pics, int skip_staged) {
    __shared__ InterWarpSmem sm[4];
    const PicDev &P = pics[blockIdx.z];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mbx = blockIdx.x * 4 + warp, mby = blockIdx.y;
    if (!P.motion || mbx >= P.wmb) return;
    inter_mb_ldg(P, mbx, mby, lane, sm[warp].raw, sm[warp].hs, skip_staged);
}
