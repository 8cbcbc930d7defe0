// common.cuh — device-side picture descriptor, geometry and neighbour derivation shared by all kernels.
//
// Reference citations: PB = H264PictureBase.cpp, IP = H264InterPrediction.cpp,
// DB = H264PictureDeblockingFilterProcess.cpp, MB = H264MacroBlock.cpp of jfu222/h264_video_decoder_demo.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "h264_recon_b200.h"

// One entry per picture of a submitted batch.  All pointers are device pointers.
struct alignas(16) PicDev {
    const H264B2MbInfo   *info;
    const uint64_t       *modes;
    const uint32_t       *coef_off;
    const H264B2MbMotion *motion;      // nullptr when the picture has no inter MB
    const H264B2Weight   *weights;
    const int16_t        *coefs;
    const uint32_t       *packed;      // packed coefficient blob (h264b2_pack_coefs) that k_expand turns into coefs[], or nullptr
    const uint32_t       *packed_motion; // packed motion blob (h264b2_pack_motion) that k_expand + k_unmotion turn into motion[], or nullptr
    const int16_t        *ls4;         // [2 inter][2 field scan][6][16] LevelScale4x4 in list order (PB:4852)
    const int16_t        *ls8;         // [2][2][6][64]
    uint8_t              *dst;         // Y plane of the destination surface; Cb = dst + W*H, Cr = Cb + W*H/4
    const uint8_t        *stream_base; // surface 0 of this picture's stream (reference surfaces = base + slot*frame_bytes)
    uint32_t             *bs;          // [n_mbs][64] packed boundary strengths (k_bs -> k_deblock)
    int16_t              *res;         // [n_mbs][384] residual scratch (k_residual -> k_inter / k_intra)
    int                  *progress;    // [2][hmb] wavefront progress counters: [0] intra, [1] deblock
    unsigned long long    frame_bytes;
    int wmb, hmb, mbaff, cqp0, cqp1;
    int deblock_enable, deblock_stop;
    int n_weights;
    int generic;                       // 1: MBAFF picture (or wider than 256 MBs): literal per-sample-line paths; 0: progressive fast paths
    int surf0;                         // index of the stream's surface 0 in the DPB allocation (third coordinate of the TMA tensor maps)
    int *worklist;                     // [0] = number of entries, [4 ..] = addresses of the inter macroblocks k_inter_tma leaves to k_inter_list
    int spp;                           // DPB surfaces per stream: reference codes are clamped to it (a corrupt SoA must not reach another stream's DPB)
    uint32_t n_coefs;                  // int16 elements behind coefs: macroblocks whose blocks would lie beyond it are treated as uncoded
};
static_assert(sizeof(PicDev) % 16 == 0, "PicDev must be a multiple of 16 bytes");

__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }

// ---- geometry (PB:2503-2527; MB:936) ----
__device__ __forceinline__ int mb_is_field(const PicDev &P, int a) { return (P.info[a].flags & H264B2_MBF_FIELD) ? 1 : 0; }
__device__ __forceinline__ void mb_origin(const PicDev &P, int a, int field, int &x0, int &y0) {
    if (!P.mbaff) { x0 = (a % P.wmb) * 16; y0 = (a / P.wmb) * 16; return; }
    int pair = a >> 1;
    x0 = (pair % P.wmb) * 16;
    int yb = (pair / P.wmb) * 32;
    y0 = field ? yb + (a & 1) : yb + (a & 1) * 16;
}
__device__ __forceinline__ int chroma_y0(int y0) { return (y0 >> 4) * 8 + (y0 & 1); }   // PB:2133
__device__ __forceinline__ int blk_x(int b) { return ((b >> 2) & 1) * 8 + (b & 1) * 4; }  // 6.4.3
__device__ __forceinline__ int blk_y(int b) { return (b >> 3) * 8 + ((b >> 1) & 1) * 4; }

// ---- neighbouring locations 6.4.12 (PB:2878 non-MBAFF, PB:2984 MBAFF) ----
__device__ __forceinline__ int avail_addr(const PicDev &P, int cur, int n) {
    if (n < 0 || n > cur) return 0;
    return P.info[n].slice_number == P.info[cur].slice_number;
}
__device__ inline int nbr_nonmbaff(const PicDev &P, int cur, int xN, int yN, int maxW, int maxH, int &xW, int &yW) {
    const int w = P.wmb;
    int n = -1;
    if (xN < 0 && yN < 0)                    { n = cur - w - 1; if (cur % w == 0) n = -1; }
    else if (xN < 0 && yN < maxH)            { n = cur - 1;     if (cur % w == 0) n = -1; }
    else if (xN >= 0 && xN < maxW && yN < 0) { n = cur - w; }
    else if (xN >= 0 && xN < maxW && yN >= 0 && yN < maxH) { xW = xN; yW = yN; return cur; }
    else if (xN >= maxW && yN < 0)           { n = cur - w + 1; if ((cur + 1) % w == 0) n = -1; }
    else return -1;
    if (n < 0 || !avail_addr(P, cur, n)) return -1;
    xW = (xN + maxW) % maxW; yW = (yN + maxH) % maxH;
    return n;
}
__device__ inline int nbr_mbaff(const PicDev &P, int cur, int xN, int yN, int maxW, int maxH, int &xW, int &yW) {
    const int w = P.wmb, pr = cur >> 1;
    int A = 2 * (pr - 1), B = 2 * (pr - w), C = 2 * (pr - w + 1), D = 2 * (pr - w - 1);
    if (!avail_addr(P, cur, A) || pr % w == 0) A = -2;
    if (!avail_addr(P, cur, B)) B = -2;
    if (!avail_addr(P, cur, C) || (pr + 1) % w == 0) C = -2;
    if (!avail_addr(P, cur, D) || pr % w == 0) D = -2;
    const int curFrame = !mb_is_field(P, cur), top = !(cur & 1);
    int n = -1, yM = 0;
#define XFRM(X) (!mb_is_field(P, (X)))
    if (xN < 0 && yN < 0) {
        if (curFrame) {
            if (top) { n = D + 1; yM = yN; }
            else if (A >= 0) { if (XFRM(A)) { n = A; yM = yN; } else { n = A + 1; yM = (yN + maxH) >> 1; } }
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
        xW = xN; yW = yN; return cur;
    } else if (xN >= maxW && yN < 0) {
        if (curFrame) { if (top) { n = C + 1; yM = yN; } else n = -1; }
        else { if (top) { if (C >= 0) { if (XFRM(C)) { n = C + 1; yM = 2 * yN; } else { n = C; yM = yN; } } } else { n = C + 1; yM = yN; } }
    }
#undef XFRM
    if (n < 0) return -1;
    xW = (xN + maxW) % maxW; yW = (yM + maxH) % maxH;
    return n;
}
__device__ __forceinline__ int nbr_loc(const PicDev &P, int cur, int xN, int yN, int chroma, int &xW, int &yW) {
    const int m = chroma ? 8 : 16;
    return P.mbaff ? nbr_mbaff(P, cur, xN, yN, m, m, xW, yW) : nbr_nonmbaff(P, cur, xN, yN, m, m, xW, yW);
}

// constructed (pre-deblock) sample of the CURRENT picture at a neighbouring location, or -1 when the
// location is not available for intra prediction (PB:1128-1154, 1492-1514, 1905-1928, 2158-2186).
// Reads bypass L1 (ld.global.cg): the sample may have been written by another SM in this launch.
__device__ inline int nbr_sample(const PicDev &P, int cur, int xN, int yN, int comp) {
    int xW, yW;
    const int n = nbr_loc(P, cur, xN, yN, comp != 0, xW, yW);
    if (n < 0) return -1;
    if (P.info[n].flags & H264B2_MBF_CIP_UNAVAIL) return -1;
    const int f = P.mbaff && mb_is_field(P, n);
    int x0, y0;
    mb_origin(P, n, f, x0, y0);
    const int W = P.wmb * 16, H = P.hmb * 16;
    if (comp == 0) return __ldcg(P.dst + (size_t)(y0 + (f ? 2 * yW : yW)) * W + x0 + xW);
    const uint8_t *pl = P.dst + (size_t)W * H + (comp == 2 ? (size_t)(W / 2) * (H / 2) : 0);
    return __ldcg(pl + (size_t)(chroma_y0(y0) + (f ? 2 * yW : yW)) * (W / 2) + (x0 >> 1) + xW);
}

// ---- bounds of host-supplied indices (the pre-parsed container path reads them straight from a file) ----
// int16 elements the coefficient blocks of a macroblock occupy (include/h264_recon_b200.h: coef_mask)
__device__ __forceinline__ uint32_t mb_coef_count(uint32_t m, int cls, int t8flag) {
    if (m & H264B2_CM_PCM) return 384u;
    const int t8 = t8flag && cls != H264B2_MB_I16x16;
    return (uint32_t)(__popc(m & 0xFFFFu) * (t8 ? 64 : 16) + ((m >> 16) & 1u) * 16 + ((m >> 17) & 1u) * 8 + __popc((m >> 18) & 0xFFu) * 16);
}
__device__ __forceinline__ bool mb_coefs_in_bounds(const PicDev &P, int a, uint32_t m, int cls, int t8flag) {
    const unsigned long long end = (unsigned long long)P.coef_off[a] + mb_coef_count(m, cls, t8flag);
    return end <= (unsigned long long)P.n_coefs;
}

// ---- chroma QP (PB:4748; table PB:4773) ----
__device__ const uint8_t g_qpc_tab[22] = {29,30,31,32,32,33,34,34,35,35,36,36,37,37,37,38,38,38,39,39,39,39};
__device__ __forceinline__ int chroma_qp(const PicDev &P, int qpy, int c) {
    const int qpi = clip3i(0, 51, qpy + (c ? P.cqp1 : P.cqp0));
    return qpi < 30 ? qpi : g_qpc_tab[qpi - 30];
}
