// residual.cuh — dequantisation + inverse integer transforms.
//
// Reference: scaling + 4x4 transform PB:4106-4266, 8x8 PB:4270-4402, Intra16x16 DC PB:4993-5088,
// chroma DC PB:3989-4060, inverse scans PB:4542-4744, LevelScale PB:4852-4989, drivers PB:3401-3929
// and IP:22-407.  Coefficients arrive in LIST order (before inverse scan) as int16; the LevelScale
// tables are pre-permuted into list order too, so dequantisation happens before the scatter.
#pragma once
#include "common.cuh"

// scan tables: [0] zig-zag, [1] field scan; forward = list index k -> raster position, inverse = position -> k
__constant__ uint8_t c_iscan4[2][16];
__constant__ uint8_t c_iscan8[2][64];

struct ResidualTile {
    int16_t res[384];      // luma 16x16 raster, then Cb 8x8, then Cr 8x8
    int     dcY[16];       // Intra16x16 luma DC after the Hadamard, raster [i][j] of 4x4 blocks
    int     dcC[2][4];     // chroma DC after the 2x2 Hadamard
};

// Residuals are stored as int16.  The reference keeps them in int32 and only ever uses them in
// Clip1(pred + r) with pred in [0,255], so saturating r to +-4096 cannot change any output sample.
__device__ __forceinline__ int16_t sat_res(int v) { return (int16_t)min(max(v, -4096), 4096); }

// 4x4 block owned by one thread.  lv: 16 list-order levels or nullptr; dc_pass: position 0 takes dcval as is.
__device__ inline void resid4x4_thread(const int16_t *lv, int dcval, int dc_pass, int qp, const int16_t *ls /*[6][16]*/,
                                       int field, int16_t *out, int ostride) {
    int d[16];
    const int16_t *lsr = ls + (qp % 6) * 16;
    const int qd = qp / 6;
#pragma unroll
    for (int pos = 0; pos < 16; pos++) {
        const int k = c_iscan4[field][pos];
        int c = lv ? (int)lv[k] : 0;
        int v;
        if (qp >= 24) v = (c * (int)lsr[k]) << (qd - 4);
        else v = (c * (int)lsr[k] + (1 << (3 - qd))) >> (4 - qd);
        d[pos] = v;
    }
    if (dc_pass) d[0] = dcval;
    int f[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s0 = d[4*i], s1 = d[4*i+1], s2 = d[4*i+2], s3 = d[4*i+3];
        const int e0 = s0 + s2, e1 = s0 - s2, e2 = (s1 >> 1) - s3, e3 = s1 + (s3 >> 1);
        f[4*i] = e0 + e3; f[4*i+1] = e1 + e2; f[4*i+2] = e1 - e2; f[4*i+3] = e0 - e3;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int g0 = f[j] + f[8+j], g1 = f[j] - f[8+j], g2 = (f[4+j] >> 1) - f[12+j], g3 = f[4+j] + (f[12+j] >> 1);
        out[0*ostride + j] = sat_res((g0 + g3 + 32) >> 6);
        out[1*ostride + j] = sat_res((g1 + g2 + 32) >> 6);
        out[2*ostride + j] = sat_res((g1 - g2 + 32) >> 6);
        out[3*ostride + j] = sat_res((g0 - g3 + 32) >> 6);
    }
}

__device__ __forceinline__ void butterfly8(int &a0, int &a1, int &a2, int &a3, int &a4, int &a5, int &a6, int &a7) {   // PB:4332-4390
    const int e0 = a0 + a4, e1 = -a3 + a5 - a7 - (a7 >> 1), e2 = a0 - a4, e3 = a1 + a7 - a3 - (a3 >> 1);
    const int e4 = (a2 >> 1) - a6, e5 = -a1 + a7 + a5 + (a5 >> 1), e6 = a2 + (a6 >> 1), e7 = a3 + a5 + a1 + (a1 >> 1);
    const int f0 = e0 + e6, f1 = e1 + (e7 >> 2), f2 = e2 + e4, f3 = e3 + (e5 >> 2), f4 = e2 - e4, f5 = (e3 >> 2) - e5, f6 = e0 - e6, f7 = e7 - (e1 >> 2);
    a0 = f0 + f7; a1 = f2 + f5; a2 = f4 + f3; a3 = f6 + f1; a4 = f6 - f1; a5 = f4 - f3; a6 = f2 - f5; a7 = f0 - f7;
}

// 8x8 block owned by one thread (lv non-null).
__device__ inline void resid8x8_thread(const int16_t *lv, int qp, const int16_t *ls /*[6][64]*/, int field, int16_t *out, int ostride) {
    int d[64];
    const int16_t *lsr = ls + (qp % 6) * 64;
    const int qd = qp / 6;
#pragma unroll
    for (int pos = 0; pos < 64; pos++) {
        const int k = c_iscan8[field][pos];
        const int c = lv[k];
        d[pos] = qp >= 36 ? (c * (int)lsr[k]) << (qd - 6) : (c * (int)lsr[k] + (1 << (5 - qd))) >> (6 - qd);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) butterfly8(d[8*i], d[8*i+1], d[8*i+2], d[8*i+3], d[8*i+4], d[8*i+5], d[8*i+6], d[8*i+7]);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        butterfly8(d[j], d[8+j], d[16+j], d[24+j], d[32+j], d[40+j], d[48+j], d[56+j]);
#pragma unroll
        for (int i = 0; i < 8; i++) out[i * ostride + j] = sat_res((d[8*i+j] + 32) >> 6);
    }
}

__device__ inline void luma_dc16_thread(const int16_t *lv, int qp, int ls00, int field, int *dcY) {   // PB:4993
    int c[16];
#pragma unroll
    for (int pos = 0; pos < 16; pos++) c[pos] = lv[c_iscan4[field][pos]];
    int g[16];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int a = c[j], b = c[4+j], cc = c[8+j], d = c[12+j];
        g[j] = a + b + cc + d; g[4+j] = a + b - cc - d; g[8+j] = a - b - cc + d; g[12+j] = a - b + cc - d;
    }
    const int qd = qp / 6;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int a = g[4*i], b = g[4*i+1], cc = g[4*i+2], d = g[4*i+3];
        const int f0 = a + b + cc + d, f1 = a + b - cc - d, f2 = a - b - cc + d, f3 = a - b + cc - d;
        const int ff[4] = { f0, f1, f2, f3 };
#pragma unroll
        for (int j = 0; j < 4; j++)
            dcY[4*i+j] = qp >= 36 ? (ff[j] * ls00) << (qd - 6) : (ff[j] * ls00 + (1 << (5 - qd))) >> (6 - qd);
    }
}

struct SyncWarp { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
struct SyncCta  { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// Compute the whole residual of macroblock `a` into T (zero where nothing is coded).
// tid/nthreads: the cooperating group; sync() must be a barrier for exactly that group.
template <class Sync>
__device__ inline void mb_residual(const PicDev &P, int a, const H264B2MbInfo &I, int tid, int nthreads, ResidualTile &T, Sync sync) {
    const int cls = I.mb_class;
    const uint32_t m = mb_coefs_in_bounds(P, a, I.coef_mask, cls, I.flags & H264B2_MBF_T8x8) ? I.coef_mask : 0u;
    const int inter = cls == H264B2_MB_INTER;
    const int sf = (I.flags & H264B2_MBF_FIELD) ? 1 : 0;     // field_pic_flag | mb_field_decoding_flag (PB:3419)
    const int t8 = (I.flags & H264B2_MBF_T8x8) && cls != H264B2_MB_I16x16;
    const int qp = I.qpy;
    const int16_t *q = P.coefs + P.coef_off[a];
    const int16_t *pldc = q + __popc(m & 0xFFFFu) * (t8 ? 64 : 16);
    const int16_t *pcdc = pldc + ((m >> 16) & 1) * 16;
    const int16_t *pcb = pcdc + ((m >> 17) & 1) * 8;
    const int16_t *pcr = pcb + __popc((m >> 18) & 15u) * 16;
    const int16_t *ls4 = P.ls4 + ((inter * 2 + sf) * 6) * 16;   // chroma uses the luma list of the same MB (Q7)
    const int16_t *ls8 = P.ls8 + ((inter * 2 + sf) * 6) * 64;
    if (tid == 0) {
        if (cls == H264B2_MB_I16x16 && (m & H264B2_CM_LUMA_DC)) luma_dc16_thread(pldc, qp, ls4[(qp % 6) * 16], sf, T.dcY);
        else for (int i = 0; i < 16; i++) T.dcY[i] = 0;
    } else if (tid == 1 || tid == 2) {
        const int c = tid - 1;
        if (m & H264B2_CM_CHROMA_DC) {                       // PB:3989
            const int16_t *s = pcdc + 4 * c;
            const int qpc = chroma_qp(P, qp, c);
            const int e00 = s[0] + s[2], e01 = s[1] + s[3], e10 = s[0] - s[2], e11 = s[1] - s[3];
            const int l0 = ls4[(qpc % 6) * 16], sh = qpc / 6;
            T.dcC[c][0] = (((e00 + e01) * l0) << sh) >> 5;
            T.dcC[c][1] = (((e00 - e01) * l0) << sh) >> 5;
            T.dcC[c][2] = (((e10 + e11) * l0) << sh) >> 5;
            T.dcC[c][3] = (((e10 - e11) * l0) << sh) >> 5;
        } else { T.dcC[c][0] = T.dcC[c][1] = T.dcC[c][2] = T.dcC[c][3] = 0; }
    }
    sync();
    for (int task = tid; task < 24; task += nthreads) {
        if (task < 16) {
            const int b = task;
            if (t8) {
                if (b < 4) {
                    int16_t *o = T.res + ((b >> 1) * 8) * 16 + (b & 1) * 8;
                    if ((m >> b) & 1) resid8x8_thread(q + 64 * __popc(m & ((1u << b) - 1)), qp, ls8, sf, o, 16);
                    else for (int i = 0; i < 64; i++) o[(i >> 3) * 16 + (i & 7)] = 0;
                }
            } else {
                int16_t *o = T.res + blk_y(b) * 16 + blk_x(b);
                const int16_t *lv = ((m >> b) & 1) ? q + 16 * __popc(m & ((1u << b) - 1)) : nullptr;
                const int is16 = cls == H264B2_MB_I16x16;
                const int dc = is16 ? T.dcY[(blk_y(b) >> 2) * 4 + (blk_x(b) >> 2)] : 0;
                if (lv || dc) resid4x4_thread(lv, dc, is16, qp, ls4, sf, o, 16);
                else for (int i = 0; i < 16; i++) o[(i >> 2) * 16 + (i & 3)] = 0;
            }
        } else {
            const int c = (task - 16) >> 2, b = (task - 16) & 3;
            int16_t *o = T.res + 256 + c * 64 + (b >> 1) * 4 * 8 + (b & 1) * 4;
            const uint32_t bit = c ? H264B2_CM_CR(b) : H264B2_CM_CB(b);
            const uint32_t below = (c ? (m >> 22) : (m >> 18)) & ((1u << b) - 1);
            const int16_t *lv = (m & bit) ? (c ? pcr : pcb) + 16 * __popc(below) : nullptr;
            const int dc = T.dcC[c][b];
            if (lv || dc) resid4x4_thread(lv, dc, 1, chroma_qp(P, qp, c), ls4, sf, o, 8);
            else for (int i = 0; i < 16; i++) o[(i >> 2) * 8 + (i & 3)] = 0;
        }
    }
    sync();
}
