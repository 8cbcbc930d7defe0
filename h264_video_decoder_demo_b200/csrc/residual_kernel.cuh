// residual_kernel.cuh — k_residual: dequantisation + inverse transforms for every coded macroblock of the
// batch, fully parallel, BEFORE prediction.  Output: int16 residual tiles in an HBM scratch that the
// prediction kernels (k_inter, k_intra) add to their prediction (T7/T8: u = Clip1(pred + r)).
//
// Scratch layout per macroblock: 24 blocks of 16 int16 (768 B):
//   blocks 0..15  luma 4x4 blocks in RASTER order inside the MB (slot = (y/4)*4 + x/4), samples [y%4][x%4]
//   blocks 16..19 Cb 4x4 blocks (raster 2x2), blocks 20..23 Cr.
// Only blocks for which res_block_coded() is true are written (and later read).
//
// Mapping: one warp per macroblock.  4x4 transforms: one lane per block, everything in registers.
// 8x8 transforms: 8 lanes per block — lane i dequantises and transforms row i, the 8x8 tile is transposed
// across the 8 lanes with shuffles, lane i transforms column i (PB:4332-4390 order: rows first).
#pragma once
#include "common.cuh"
#include "residual.cuh"

#define RES_MB_STRIDE 384   // int16 per macroblock

// does raster luma slot r / chroma block (c,b) of this MB carry a residual?
__device__ __forceinline__ int luma_slot_coded(uint32_t mask, int cls, int t8eff, int r) {
    const int bxq = (r & 3) >> 1, byq = r >> 3;              // quadrant
    if (t8eff) return (mask >> (byq * 2 + bxq)) & 1;
    const int b = byq * 8 + bxq * 4 + ((r >> 2) & 1) * 2 + (r & 1);   // luma4x4BlkIdx of raster slot r (inverse of 6.4.3)
    return ((mask >> b) & 1) | (cls == H264B2_MB_I16x16 ? (mask >> 16) & 1 : 0);
}
__device__ __forceinline__ int chroma_blk_coded(uint32_t mask, int c, int b) {
    return ((mask >> 17) & 1) | ((mask >> ((c ? 22 : 18) + b)) & 1);
}
__device__ __forceinline__ int mb_has_residual(const H264B2MbInfo &I) {
    return I.mb_class != H264B2_MB_NA && I.mb_class != H264B2_MB_IPCM && (I.coef_mask & 0x3FFFFFFu) != 0;
}

__device__ __forceinline__ void store_blk16(int16_t *dst, const int16_t *v) {
    uint4 a, b;
    a.x = (uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);  a.y = (uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
    a.z = (uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);  a.w = (uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
    b.x = (uint16_t)v[8] | ((uint32_t)(uint16_t)v[9] << 16);  b.y = (uint16_t)v[10] | ((uint32_t)(uint16_t)v[11] << 16);
    b.z = (uint16_t)v[12] | ((uint32_t)(uint16_t)v[13] << 16); b.w = (uint16_t)v[14] | ((uint32_t)(uint16_t)v[15] << 16);
    ((uint4 *)dst)[0] = a; ((uint4 *)dst)[1] = b;
}

__global__ void __launch_bounds__(256) k_residual(const PicDev *pics) {
    const PicDev &P = pics[blockIdx.y];
    const int a = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (a >= P.wmb * P.hmb) return;
    const H264B2MbInfo I = P.info[a];
    if (!mb_has_residual(I)) return;
    const int cls = I.mb_class;
    // blocks that would lie beyond the coefficient array (corrupt offsets) are treated as uncoded
    const uint32_t m = mb_coefs_in_bounds(P, a, I.coef_mask, cls, I.flags & H264B2_MBF_T8x8) ? I.coef_mask : 0u;
    if (m == 0u && cls == H264B2_MB_INTER) return;
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
    int16_t *out = P.res + (size_t)a * RES_MB_STRIDE;
    // intra MBs get ALL 24 blocks written (zeros where nothing is coded): k_intra stages the whole tile with
    // plain vector loads; inter MBs only get their coded blocks (k_inter tests the mask per block)
    const int full = !inter;

    if (t8) {
        // ---- luma 8x8: 8 lanes per block
        const int b = lane >> 3, i = lane & 7;
        const int coded = (m >> b) & 1;                     // uniform across the 8 lanes of a block
        int d[8];
        if (coded) {
            const int16_t *lv = q + 64 * __popc(m & ((1u << b) - 1));
            const int16_t *lsr = ls8 + (qp % 6) * 64;
            const int qd = qp / 6;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = c_iscan8[sf][i * 8 + j];
                const int c = lv[k];
                d[j] = qp >= 36 ? (c * (int)lsr[k]) << (qd - 6) : (c * (int)lsr[k] + (1 << (5 - qd))) >> (6 - qd);
            }
            butterfly8(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) d[j] = 0;
        }
        // transpose the 8x8 tile across the 8 lanes of the block (3 xor-butterfly stages):
        // afterwards d[j] = element (row j, column i)
#pragma unroll
        for (int k = 1; k < 8; k <<= 1) {
            const bool up = (lane & k) != 0;
#pragma unroll
            for (int j0 = 0; j0 < 8; j0++) {
                if (j0 & k) continue;
                const int mine = up ? d[j0] : d[j0 | k];
                const int got = __shfl_xor_sync(0xffffffffu, mine, k);
                if (up) d[j0] = got; else d[j0 | k] = got;
            }
        }
        if (coded || full) {
            butterfly8(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
            // lane i owns column i of 8x8 block b: scatter into the raster 4x4 slots
            const int xq = (b & 1) * 8 + i, yq = (b >> 1) * 8;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int x = xq, y = yq + j;
                out[((y >> 2) * 4 + (x >> 2)) * 16 + (y & 3) * 4 + (x & 3)] = sat_res((d[j] + 32) >> 6);
            }
        }
    }
    // ---- 4x4 blocks: luma lanes 0..15 (lane = luma4x4BlkIdx; not with the 8x8 transform) and chroma lanes 16..23 (c = Cb/Cr, b = block) run
    //      ONE copy of the 4x4 dequantisation + transform together (they used to be two divergent sections, executed one after the other)
    const int16_t *lv4 = nullptr;
    int16_t *dst4 = nullptr;
    int dc4 = 0, dcf4 = 0, qp4 = qp;
    bool do4 = false;
    if (!t8 && lane < 16) {
        const int b = lane;
        const int is16 = cls == H264B2_MB_I16x16;
        lv4 = ((m >> b) & 1) ? q + 16 * __popc(m & ((1u << b) - 1)) : nullptr;
        if (is16 && (m & H264B2_CM_LUMA_DC)) {
            int dcY[16];
            luma_dc16_thread(pldc, qp, ls4[(qp % 6) * 16], sf, dcY);     // redundant per lane, no communication
            const int idx = (blk_y(b) >> 2) * 4 + (blk_x(b) >> 2);
#pragma unroll
            for (int j = 0; j < 16; j++) if (j == idx) dc4 = dcY[j];
        }
        dcf4 = is16;
        do4 = lv4 || (is16 && (m & H264B2_CM_LUMA_DC)) || full;
        dst4 = out + ((blk_y(b) >> 2) * 4 + (blk_x(b) >> 2)) * 16;
    } else if (lane >= 16 && lane < 24) {
        const int c = (lane - 16) >> 2, b = (lane - 16) & 3;
        if (chroma_blk_coded(m, c, b) || full) {
            qp4 = chroma_qp(P, qp, c);
            if (m & H264B2_CM_CHROMA_DC) {                   // PB:3989
                const int16_t *s = pcdc + 4 * c;
                const int e00 = s[0] + s[2], e01 = s[1] + s[3], e10 = s[0] - s[2], e11 = s[1] - s[3];
                const int f = b == 0 ? e00 + e01 : b == 1 ? e00 - e01 : b == 2 ? e10 + e11 : e10 - e11;
                dc4 = ((f * (int)ls4[(qp4 % 6) * 16]) << (qp4 / 6)) >> 5;
            }
            const uint32_t below = (c ? (m >> 22) : (m >> 18)) & ((1u << b) - 1);
            lv4 = ((m >> ((c ? 22 : 18) + b)) & 1) ? (c ? pcr : pcb) + 16 * __popc(below) : nullptr;
            dcf4 = 1; do4 = true;
            dst4 = out + (16 + c * 4 + b) * 16;
        }
    }
    if (do4) {
        int16_t r[16];
        resid4x4_thread(lv4, dc4, dcf4, qp4, ls4, sf, r, 4);
        store_blk16(dst4, r);
    }
}
