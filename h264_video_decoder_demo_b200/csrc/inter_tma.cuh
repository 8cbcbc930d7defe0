// inter_tma.cuh — k_inter_tma: inter prediction + inter residual add with the reference windows staged into shared memory by
// the Tensor Memory Accelerator.
//
// Reference: Inter_prediction_process IP:412-667, fractional sample interpolation IP:2228-2328 (36 clamped loads per sample,
// IP:2344-2480), chroma IP:2485-2522, weighting IP:2526-2829, residual add IP:22-407.
//
// Round 1's k_inter (inter_quad.cuh) pulled every window row with four aligned LDG.32 + funnel shifts per lane and was held
// by L2 latency (ncu: long-scoreboard its top stall) and by ~300 instructions of per-lane address arithmetic per macroblock.
// Here the decoded picture buffer is described ONCE by two tensor maps (engine.cu):
//     luma    (x, y, surface)          u8, box 48 x 21      — the 21 x 21 window of a 16x16 partition
//     chroma  (x, y, plane, surface)   u8, box 32 x 9 x 2   — the 9 x 9 windows of Cb and Cr in one transfer
// (TMA wants the box to start on a 16-byte address in the innermost dimension — an unaligned x coordinate is an illegal
// instruction, tools/tma_probe.cu — so the box starts at the window's x rounded down to 16 and is 15 bytes wider)
// and a warp walks IT_MBS consecutive macroblocks as a two-stage pipeline: while it filters macroblock i out of shared
// memory, one lane has already issued cp.async.bulk.tensor for macroblock i+1 into the other buffer (completion on an
// mbarrier, expect_tx = the box bytes).  No lane computes a global address for a reference sample and no load waits on L2 in
// the filter.  The filter itself is the quadrant mapping of inter_quad.cuh (8 lanes per 8x8 quadrant, DP4A horizontal sums,
// packed 2 x 16-bit vertical taps): each lane realigns its two window rows out of the staged tile (4 LDS + funnel shifts).
// What takes this path: frame macroblocks of progressive pictures that carry ONE vector per list (81 % of the inter
// macroblocks of the bundled streams) and whose windows lie inside the picture.  TMA fills out-of-bounds elements with
// zeros, the reference clamps coordinates (IP:2363) — so windows that leave the picture, macroblocks with several vectors,
// and MBAFF pictures keep the clamped / per-quadrant routine of inter_quad.cuh (inter_mb_ldg below), bit-exact as before.
#pragma once
#include <cuda.h>
#include "inter_quad.cuh"
#include "wavefront.cuh"

#ifndef IT_MBS
#define IT_MBS 5                     // macroblocks per warp (one after the other)
#endif
#define IT_LUMA_PITCH 48
#define IT_LUMA_ROWS 21
#define IT_CHROMA_PITCH 32
#define IT_CHROMA_ROWS 9
#define IT_LUMA_BYTES 1024           // 21 x 48 = 1008, padded to the 128-byte alignment a TMA destination needs
#define IT_CHROMA_BYTES 640          // 2 x 9 x 32 = 576, padded
#define IT_LIST_BYTES (IT_LUMA_BYTES + IT_CHROMA_BYTES)
#define IT_TX_BYTES (IT_LUMA_PITCH * IT_LUMA_ROWS + 2 * IT_CHROMA_PITCH * IT_CHROMA_ROWS)      // bytes one list's two boxes deliver


__device__ __forceinline__ void tma_load_3d(uint32_t dst, unsigned long long map, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, unsigned long long map, int x, int y, int p, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(p), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" :: "r"(bar), "r"(bytes) : "memory");
}

// Predicted luma row r of one quadrant out of a staged box.  win: byte offset, inside the box in shared memory, of the quadrant's
// window origin (sample (xI - 2, yI - 2)); any alignment.  Same arithmetic as luma_quad_pred, rows come from shared memory.
__device__ __forceinline__ void luma_quad_pred_smem(const uint8_t *box, int win, int xF, int yF, int r, unsigned gmask, uint32_t (*raw)[4], int (*hs)[8], uint32_t out[2]) {
    const bool needJ = (xF == 2 && yF != 0) || (yF == 2 && xF != 0);
    const int ry = 2 + (yF == 3);
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int w = r + 8 * t;
        if (w < 13 && (yF != 0 || (w >= 2 && w < 10))) {
            const int off = win + w * IT_LUMA_PITCH;
            const uint32_t *pw = (const uint32_t *)(box + (off & ~3));
            const int sh = (off & 3) * 8;
            const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2], w3 = pw[3];
            const uint32_t A0 = __funnelshift_r(w0, w1, sh), A1 = __funnelshift_r(w1, w2, sh), A2 = __funnelshift_r(w2, w3, sh), A3 = w3 >> sh;
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

// Row cy of the quadrant's 4x4 chroma block of one plane out of a staged box (IP:2485-2522).  win: byte offset of sample (xC, yC + cy).
__device__ __forceinline__ void chroma_quad_pred_smem(const uint8_t *box, int win, int xF, int yF, int p[4]) {
    uint32_t s[2], e[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int off = win + k * IT_CHROMA_PITCH;
        const uint32_t *pw = (const uint32_t *)(box + (off & ~3));
        const int sh = (off & 3) * 8;
        const uint32_t w0 = pw[0], w1 = pw[1];
        s[k] = __funnelshift_r(w0, w1, sh); e[k] = (w1 >> sh) & 0xffu;
    }
    const uint32_t cf = (uint32_t)((8 - xF) * (8 - yF)) | ((uint32_t)(xF * (8 - yF)) << 8) | ((uint32_t)((8 - xF) * yF) << 16) | ((uint32_t)(xF * yF) << 24);
    const uint32_t h0 = (s[0] >> 24) | (e[0] << 8), h1 = (s[1] >> 24) | (e[1] << 8);
    p[0] = dp4a_uu(__byte_perm(s[0], s[1], 0x5410), cf, 32) >> 6;
    p[1] = dp4a_uu(__byte_perm(s[0], s[1], 0x6521), cf, 32) >> 6;
    p[2] = dp4a_uu(__byte_perm(s[0], s[1], 0x7632), cf, 32) >> 6;
    p[3] = dp4a_uu(__byte_perm(h0, h1, 0x5410), cf, 32) >> 6;
}

// The macroblocks k_inter_tma does not take (several vectors per list, windows that leave the picture), from its work list: every
// warp takes entries warp, warp + n_warps, ... and runs the clamped / per-quadrant routine of inter_quad.cuh on them.
// grid: (IT_LIST_CTAS, n_pics), block 128.
#ifndef IT_LIST_CTAS
#define IT_LIST_CTAS 32
#endif
__global__ void __launch_bounds__(128, INTER_MIN_BLOCKS) k_inter_list(const PicDev *pics) {
    __shared__ InterWarpSmem sm[4];
    const PicDev &P = pics[blockIdx.y];
    if (!P.motion) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int count = P.worklist[0], wmb = P.wmb;
    for (int e = blockIdx.x * 4 + warp; e < count; e += IT_LIST_CTAS * 4) {
        const int a = P.worklist[4 + e];
        const int mby = a / wmb;
        inter_mb_ldg(P, a - mby * wmb, mby, lane, sm[warp].raw, sm[warp].hs, 0);
        __syncwarp();
    }
}
__global__ void k_zero_worklists(int *w, int n, size_t stride) { for (int i = threadIdx.x; i < n; i += blockDim.x) w[(size_t)i * stride] = 0; }

// what the set-up of a macroblock leaves for its filter stage (per warp, in shared memory)
struct ItMb {
    int mode;                  // 1: staged by TMA and reconstructed here; 0: not this kernel's macroblock (k_inter does the others)
    uint32_t mv0, mv1;         // the one vector of each list
    int code0, code1, wt;
    uint32_t coef_mask, flags;
};
struct alignas(128) InterTmaWarp {
    uint8_t tile[2][2][IT_LIST_BYTES];        // [pipeline stage][list]: luma box, then the Cb | Cr box
    int hs[4][13][8];                          // unclipped horizontal 6-tap sums per quadrant
    uint32_t raw[4][13][4];                    // realigned window rows per quadrant
    ItMb mb[IT_MBS];
    unsigned long long bar[2];
};

// grid: (ceil(wmb / (4 * IT_MBS)), hmb, n_pics); block: 128 threads = 4 warps, warp w walks macroblocks
// (blockIdx.x * 4 + w) * IT_MBS .. + IT_MBS - 1 of row blockIdx.y.  Launched only for progressive pictures (PicDev::generic == 0).
#ifndef IT_MIN_BLOCKS
#define IT_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(128, IT_MIN_BLOCKS) k_inter_tma(const PicDev *pics, const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapC) {
    __shared__ InterTmaWarp sm[4];
    const PicDev &P = pics[blockIdx.z];
    if (!P.motion || P.mbaff || P.generic) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    InterTmaWarp &S = sm[warp];
    // addresses of the descriptors in the kernel parameter space
    const unsigned long long map_y = (unsigned long long)&mapY, map_c = (unsigned long long)&mapC;
    const int wmb = P.wmb, mby = blockIdx.y;
    const int mb0 = (blockIdx.x * 4 + warp) * IT_MBS;
    if (mb0 >= wmb) return;
    const int nmb_w = min(IT_MBS, wmb - mb0);
    const uint32_t bar0 = smem_addr(&S.bar[0]);
    if (lane == 0) { mbar_init((uint64_t *)&S.bar[0], 1); mbar_init((uint64_t *)&S.bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    const int q = lane >> 3, r = lane & 7, qx = (q & 1) * 8, qy = (q >> 1) * 8;
    const unsigned gmask = 0xFFu << (q * 8);
    const int cpl = r >> 2, cy = r & 3;

    // ---- phase A: side information of all macroblocks of this warp.  Every load is issued before the first one is used, so the
    //      warp pays ONE memory round trip for its IT_MBS macroblocks (the first version loaded and decided macroblock by
    //      macroblock and spent as long waiting for these few words as filtering).
    {
        uint4 iw[IT_MBS]; uint2 sf[IT_MBS], wt[IT_MBS]; uint32_t word[IT_MBS];
#pragma unroll
        for (int i = 0; i < IT_MBS; i++) {
            const int a = mby * wmb + min(mb0 + i, wmb - 1);
            const uint32_t *mw = (const uint32_t *)(P.motion + a);
            iw[i] = __ldg((const uint4 *)(P.info + a));
            sf[i] = __ldg((const uint2 *)mw + 16); wt[i] = __ldg((const uint2 *)mw + 18);
            word[i] = __ldg(mw + lane);                                          // lane l: vector l & 15 of list l >> 4
        }
#pragma unroll
        for (int i = 0; i < IT_MBS; i++) {
            const uint32_t first = __shfl_sync(0xffffffffu, word[i], lane & 16);
            const bool used = (int8_t)(((lane & 16) ? sf[i].y : sf[i].x) & 0xffu) >= 0;
            const bool uni = __all_sync(0xffffffffu, word[i] == first || !used) && sf[i].x == (sf[i].x & 0xffu) * 0x01010101u && sf[i].y == (sf[i].y & 0xffu) * 0x01010101u &&
                             wt[i].x == wt[i].y && (wt[i].x & 0xffffu) == (wt[i].x >> 16);
            const uint32_t mv0 = __shfl_sync(0xffffffffu, word[i], 0), mv1 = __shfl_sync(0xffffffffu, word[i], 16);
            const int code0 = (int)(int8_t)(sf[i].x & 0xffu), code1 = (int)(int8_t)(sf[i].y & 0xffu);
            const bool mine = i < nmb_w && (iw[i].x & 0xffu) == H264B2_MB_INTER && !((iw[i].x >> 8) & H264B2_MBF_FIELD) && uni && inter_staged_ok(P, mb0 + i, mby, code0, code1, mv0, mv1);
            if (lane == 0) {
                // inter macroblocks that do not qualify go on the picture's work list for k_inter_list
                if (i < nmb_w && (iw[i].x & 0xffu) == H264B2_MB_INTER && !mine) { const int e = atomicAdd(P.worklist, 1); P.worklist[4 + e] = mby * wmb + mb0 + i; }
                ItMb &m = S.mb[i];
                m.mode = mine; m.mv0 = mv0; m.mv1 = mv1; m.code0 = code0; m.code1 = code1; m.wt = (int)(wt[i].x & 0xffffu); m.coef_mask = iw[i].w; m.flags = (iw[i].x >> 8) & 0xffu;
            }
        }
    }
    __syncwarp();
    unsigned uses = 0;                                           // bit s: the parity to wait for on pipeline stage s

    // ---- TMA issue for macroblock i into stage (i & 1): one lane, two boxes per list
    auto issue = [&](int i) {
        const ItMb &m = S.mb[i];
        if (!m.mode || lane != 0) return;
        const int st = i & 1, x0 = (mb0 + i) * 16, y0 = mby * 16;
        const uint32_t bar = bar0 + 8 * st;
        mbar_expect_tx(bar, (uint32_t)IT_TX_BYTES * ((m.code0 >= 0) + (m.code1 >= 0)));
#pragma unroll
        for (int l = 0; l < 2; l++) {
            const int code = l ? m.code1 : m.code0;
            if (code < 0) continue;
            const uint32_t mv = l ? m.mv1 : m.mv0;
            const int mvx = (int16_t)(mv & 0xffffu), mvy = (int16_t)(mv >> 16);
            const uint32_t dst = smem_addr(&S.tile[st][l][0]);
            const int surf = P.surf0 + min(code >> 2, P.spp - 1);
            tma_load_3d(dst, map_y, (x0 + (mvx >> 2) - 2) & ~15, y0 + (mvy >> 2) - 2, surf, bar);
            tma_load_4d(dst + IT_LUMA_BYTES, map_c, ((x0 >> 1) + (mvx >> 3)) & ~15, (y0 >> 1) + (mvy >> 3), 0, surf, bar);
        }
    };

    issue(0);
    for (int i = 0; i < nmb_w; i++) {
        if (i + 1 < nmb_w) issue(i + 1);
        const ItMb cur = S.mb[i];
        if (cur.mode) {
            const int st = i & 1;
            const uint32_t bar = bar0 + 8 * st;
            // bounded: a transfer that never completes (bad descriptor, wrong byte count) must not hang the GPU
            { unsigned spins = 0; while (!mbar_try_wait(bar, (uses >> st) & 1)) { if (++spins > (1u << 22)) { if (lane == 0) printf("k_inter_tma: TMA wait timed out (pic %d row %d mb %d stage %d uses %u mode %d codes %d %d)\n", blockIdx.z, mby, mb0 + i, st, uses, cur.mode, cur.code0, cur.code1); __trap(); } } }
            uses ^= 1u << st;
            const int mbx = mb0 + i, a = mby * wmb + mbx;
            const int x0 = mbx * 16, y0 = mby * 16;
            const int have0 = cur.code0 >= 0, have1 = cur.code1 >= 0;
            uint32_t pl[2][2] = {{0, 0}, {0, 0}};
            int pc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 1
            for (int l = 0; l < 2; l++) {
                if ((l ? cur.code1 : cur.code0) < 0) continue;
                const uint32_t mv = l ? cur.mv1 : cur.mv0;
                const int mvx = (int16_t)(mv & 0xffffu), mvy = (int16_t)(mv >> 16);
                const uint8_t *tl = &S.tile[st][l][0];
                uint32_t o[2]; int tc[4];
                const int ox = (x0 + (mvx >> 2) - 2) & 15, oc = ((x0 >> 1) + (mvx >> 3)) & 15;      // where the windows start inside their 16-byte aligned boxes
                luma_quad_pred_smem(tl, qy * IT_LUMA_PITCH + qx + ox, mvx & 3, mvy & 3, r, gmask, S.raw[q], S.hs[q], o);
                chroma_quad_pred_smem(tl + IT_LUMA_BYTES, cpl * (IT_CHROMA_PITCH * IT_CHROMA_ROWS) + ((qy >> 1) + cy) * IT_CHROMA_PITCH + (qx >> 1) + oc, mvx & 7, mvy & 7, tc);
                if (l == 0) { pl[0][0] = o[0]; pl[0][1] = o[1]; pc[0][0] = tc[0]; pc[0][1] = tc[1]; pc[0][2] = tc[2]; pc[0][3] = tc[3]; }
                else        { pl[1][0] = o[0]; pl[1][1] = o[1]; pc[1][0] = tc[0]; pc[1][1] = tc[1]; pc[1][2] = tc[2]; pc[1][3] = tc[3]; }
            }
            inter_quad_store(P, a, cur.coef_mask, (cur.flags & H264B2_MBF_T8x8) != 0, x0, y0, 1, q, r, pl, pc, have0, have1, cur.wt);
        }
        __syncwarp();                                            // every lane is done with this stage's tile before it is refilled
    }
}
