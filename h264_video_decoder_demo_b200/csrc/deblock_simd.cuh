// deblock_simd.cuh — deblocking of progressive pictures, second generation.
//
// Reference: Deblocking_filter_process DB:76-635 (MB loop, edge order), Filtering_process_for_block_edges DB:639-840,
// boundary strengths DB:994-1310, thresholds DB:1314-1409, the two filters DB:1373-1522.
//
// The first generation (deblock_fast.cuh) filtered ONE sample line per lane with scalar arithmetic and spent ~1 100 warp
// instructions per macroblock (ncu, round 1): the kernels were instruction-issue bound at 5 % of the HBM roofline.  Here:
//   * k_bs_prog2: ONE THREAD per macroblock derives its 32 strengths with the macroblock's motion record in registers
//     (the 4 possible vector comparisons of an edge are packed 2 x 16-bit tests, reference identities are compared per 8x8
//     quadrant), and writes a 96-byte record the filter kernel can use without any table look-up: per (lane role, phase)
//     one word of "step codes" (tC0 | bS>0 | bS==4 per filter step) plus alpha/beta per (plane, edge kind).
//   * k_deblock2: the filter arithmetic runs on TWO sample lines per instruction (simd16.cuh: packed 2 x 16 bit, checked
//     against the scalar equations on the CPU), every lane owns a 4x4-sample block column/row, and one warp advances the same
//     macroblock row of FOUR pictures in lock-step: 8 lanes per macroblock (4 luma block rows, 2 x 2 chroma half planes).
//     All 32 lanes run the same instruction stream; steps in which no lane has a non-zero strength are skipped warp-wide.
//     The normative order is kept: per MB vertical edges 0,4,8,12 (lanes = 4-row groups, samples transposed in registers
//     with PRMT), then horizontal edges (lanes = 4-column groups, natural layout); the hand-over between the two phases is a
//     4x4-block transpose through a 400-byte shared tile per picture.  MB rows advance as the 2:1 wavefront of wavefront.cuh.
#pragma once
#include "common.cuh"
#include "wavefront.cuh"
#include "simd16.cuh"

#define DBREC_WORDS 24                  // words per macroblock record (k_bs_prog2 -> k_deblock2)
#define DB_PPW 4                        // pictures per warp in k_deblock2

// ------------------------------------------------------------------ boundary strengths, thread per macroblock
__device__ __forceinline__ uint32_t bs_nnz_raster(uint32_t m) {          // nnz bits: luma4x4BlkIdx order -> raster order
    return (m & 0xC3C3u) | ((m & 0x0C0Cu) << 2) | ((m & 0x3030u) >> 2);
}
// |ax - bx| >= 4 || |ay - by| >= 4 for packed (x, y) int16 vectors; nb3 = 3 - b per half
__device__ __forceinline__ int bs_far(uint32_t a, uint32_t nb3) {
    const uint32_t m = __vminu2(__vadd2(a, nb3), 0x00070007u);
    return ((m + 0x00010001u) & 0x00080008u) != 0u;
}
struct BsQ { uint32_t id, idsw; int s, r0, r1; };      // reference identities / prediction flags of one 8x8 quadrant
__device__ __forceinline__ BsQ bs_q(uint32_t surf0, uint32_t surf1, uint32_t id0, uint32_t id1, int qd) {
    BsQ q; const int sh = 8 * qd;
    q.r0 = (int)((id0 >> sh) & 0xffu); q.r1 = (int)((id1 >> sh) & 0xffu);
    q.id = (uint32_t)q.r0 | ((uint32_t)q.r1 << 8); q.idsw = (uint32_t)q.r1 | ((uint32_t)q.r0 << 8);
    q.s = (int)((((surf0 >> (sh + 7)) & 1u) ^ 1u) | ((((surf1 >> (sh + 7)) & 1u) ^ 1u) << 1));
    return q;
}
// the motion part of DB:1175-1310 for frame macroblocks (mixedModeEdgeFlag 0, vertical limit 4); Q5, Q6 as in derive_bs_core
__device__ __forceinline__ int bs_motion(const BsQ &p, const BsQ &q, uint32_t m0p, uint32_t m1p, uint32_t m0q, uint32_t m1q) {
    if (!((p.id == q.id || p.id == q.idsw) && __popc(p.s) == __popc(q.s))) return 1;
    const uint32_t n0 = __vadd2(~m0q, 0x00040004u), n1 = __vadd2(~m1q, 0x00040004u);
    const int F00 = bs_far(m0p, n0), F01 = bs_far(m0p, n1), F10 = bs_far(m1p, n0), F11 = bs_far(m1p, n1);
    const int key = p.s * 4 + q.s;
    if (key == 5) return F00;
    if (key == 6) return F01;
    if (key == 9) return F10;
    if (key == 10) return F11;
    if (key == 15) {
        if (p.r0 != p.r1) return ((q.r0 == p.r0) & (F00 | F11)) | ((q.r0 == p.r1) & (F10 | F01));
        return F00 | (F11 & F01) | F10;
    }
    return 0;
}
struct BsInfo { int cls, flags, qp, slice, nnz, offa, offb, idc; };
__device__ __forceinline__ BsInfo bs_info(const H264B2MbInfo *p) {
    const uint4 w = __ldg((const uint4 *)p);
    BsInfo I;
    I.cls = w.x & 0xff; I.flags = (w.x >> 8) & 0xff; I.qp = (int)(int8_t)(w.x >> 24); I.slice = w.y & 0xffff; I.nnz = w.y >> 16;
    I.offa = (int)(int8_t)(w.z & 0xff); I.offb = (int)(int8_t)((w.z >> 8) & 0xff); I.idc = (w.z >> 16) & 0xff;
    if (I.cls == H264B2_MB_IPCM) I.qp = 0;                                 // DB:898: qPp = 0 for I_PCM
    return I;
}
__device__ __forceinline__ int bs_intra(const BsInfo &I) { return (I.cls >= H264B2_MB_I4x4 && I.cls <= H264B2_MB_I16x16) || (I.flags & H264B2_MBF_SPSI); }
__device__ __forceinline__ uint32_t bs_code(unsigned long long lut, uint32_t n) { return (uint32_t)(lut >> (8 * n)) & 0xffu; }

// Record of macroblock a at bs[a * DBREC_WORDS] (lane roles: 0-3 luma 4-row / 4-column group, 4 + 2c + h: plane c, half h):
//   word 2*role     vertical-edge step codes, word 2*role + 1 horizontal-edge step codes.  A code byte = tC0 | 0x40 (bS > 0) | 0x80 (bS = 4).
//     luma role i : byte s = code of edge 4s, segment i (both sample lines of a filter call share it)
//     chroma role : bytes 0,1 = codes of chroma edge 0 for the two luma segments the half covers, bytes 2,3 = chroma edge 4 (luma edge 8)
//   word 16 + 2*plane : alpha_left | beta_left << 8 | alpha_top << 16 | beta_top << 24;  word 17 + 2*plane : alpha_internal | beta_internal << 8
// bs[n_mbs * 64 + a] = 1 when any strength of the macroblock is non-zero.
__global__ void __launch_bounds__(128) k_bs_prog2(const PicDev *pics) {
    __shared__ uint32_t s_ta[52], s_tb[52];
    if (threadIdx.x < 52) {
        s_ta[threadIdx.x] = (uint32_t)g_alpha_tab[threadIdx.x] | ((uint32_t)g_tc0_tab[0][threadIdx.x] << 8) | ((uint32_t)g_tc0_tab[1][threadIdx.x] << 13) | ((uint32_t)g_tc0_tab[2][threadIdx.x] << 18);
        s_tb[threadIdx.x] = g_beta_tab[threadIdx.x];
    }
    __syncthreads();
    const PicDev &P = pics[blockIdx.y];
    const int a = blockIdx.x * 128 + threadIdx.x;
    const int wmb = P.wmb, nmb = wmb * P.hmb;
    if (a >= nmb || !P.deblock_enable || a >= P.deblock_stop || P.generic) return;
    const int mby = a / wmb, mbx = a - mby * wmb;
    const BsInfo I = bs_info(P.info + a);
    BsInfo IA = I, IB = I;
    int A = -1, B = -1;                                                    // DB:13-69 / PB:2878: same slice, address <= current
    if (mbx > 0) { IA = bs_info(P.info + a - 1); if (IA.slice == I.slice) A = a - 1; }
    if (mby > 0) { IB = bs_info(P.info + a - wmb); if (IB.slice == I.slice) B = a - wmb; }
    const int t8 = (I.flags & H264B2_MBF_T8x8) != 0;
    const int left = !(mbx == 0 || I.idc == 1 || (I.idc == 2 && A < 0)), top = !(mby == 0 || I.idc == 1 || (I.idc == 2 && B < 0)), internal = I.idc != 1;   // DB:104-154
    const int qi = bs_intra(I), ai = bs_intra(IA), bi = bs_intra(IB);
    const uint32_t nz = bs_nnz_raster(I.nnz), nzA = bs_nnz_raster(IA.nnz), nzB = bs_nnz_raster(IB.nnz);
    const bool mot = P.motion != nullptr && !qi;
    uint32_t mv0[16], mv1[16], a0[4], a1[4], b0[4], b1[4];
    BsQ Q[4], QA[2], QB[2];
#pragma unroll
    for (int i = 0; i < 16; i++) mv0[i] = mv1[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { a0[i] = a1[i] = b0[i] = b1[i] = 0; Q[i] = bs_q(0, 0, 0, 0, 0); }
    QA[0] = QA[1] = QB[0] = QB[1] = Q[0];
    if (mot) {
        const uint2 *m = (const uint2 *)(P.motion + a);                    // records are 152 bytes: 8-byte aligned
#pragma unroll
        for (int i = 0; i < 8; i++) { const uint2 u = __ldg(m + i), v = __ldg(m + 8 + i); mv0[2 * i] = u.x; mv0[2 * i + 1] = u.y; mv1[2 * i] = v.x; mv1[2 * i + 1] = v.y; }
        const uint2 sf = __ldg(m + 16), idn = __ldg(m + 17);
#pragma unroll
        for (int qd = 0; qd < 4; qd++) Q[qd] = bs_q(sf.x, sf.y, idn.x, idn.y, qd);
        if (A >= 0 && !ai) {
            const uint32_t *w = (const uint32_t *)(P.motion + A);
#pragma unroll
            for (int s = 0; s < 4; s++) { a0[s] = __ldg(w + s * 4 + 3); a1[s] = __ldg(w + 16 + s * 4 + 3); }
            const uint2 sa = __ldg((const uint2 *)w + 16), ia = __ldg((const uint2 *)w + 17);
            QA[0] = bs_q(sa.x, sa.y, ia.x, ia.y, 1); QA[1] = bs_q(sa.x, sa.y, ia.x, ia.y, 3);
        }
        if (B >= 0 && !bi) {
            const uint2 *w = (const uint2 *)(P.motion + B);
            const uint2 u0 = __ldg(w + 6), u1 = __ldg(w + 7), v0 = __ldg(w + 14), v1 = __ldg(w + 15);
            b0[0] = u0.x; b0[1] = u0.y; b0[2] = u1.x; b0[3] = u1.y; b1[0] = v0.x; b1[1] = v0.y; b1[2] = v1.x; b1[3] = v1.y;
            const uint2 sb = __ldg(w + 16), ib = __ldg(w + 17);
            QB[0] = bs_q(sb.x, sb.y, ib.x, ib.y, 2); QB[1] = bs_q(sb.x, sb.y, ib.x, ib.y, 3);
        }
    }
    const bool has_motion = P.motion != nullptr;
    unsigned long long bv = 0, bh = 0;                                     // nibble index = edge * 4 + segment
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const bool on = e == 0 ? (dir ? top : left) : (internal && (!t8 || e == 2));
            if (!on) continue;
            const int nbr = dir ? B : A;
            const bool outer = e == 0 && nbr >= 0;
            uint32_t nib4 = 0;
            if (qi || (outer && (dir ? bi : ai))) nib4 = outer ? 0x4444u : 0x3333u;
            else {
                const uint32_t nzp = outer ? (dir ? nzB : nzA) : nz;
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const int rowq = dir ? e : s, colq = dir ? s : e;
                    const int rowp = dir ? ((e + 3) & 3) : s, colp = dir ? s : ((e + 3) & 3);
                    const int rq = rowq * 4 + colq, rp = rowp * 4 + colp;
                    uint32_t bS;
                    if (((nz >> rq) | (nzp >> rp)) & 1u) bS = 2;
                    else if (!has_motion) bS = 0;
                    else {
                        const BsQ &qq = Q[(rowq >> 1) * 2 + (colq >> 1)];
                        if (outer) bS = dir ? bs_motion(QB[s >> 1], qq, b0[s], b1[s], mv0[rq], mv1[rq]) : bs_motion(QA[s >> 1], qq, a0[s], a1[s], mv0[rq], mv1[rq]);
                        else bS = bs_motion(Q[(rowp >> 1) * 2 + (colp >> 1)], qq, mv0[rp], mv1[rp], mv0[rq], mv1[rq]);
                    }
                    nib4 |= bS << (4 * s);
                }
            }
            if (dir) bh |= (unsigned long long)nib4 << (16 * e); else bv |= (unsigned long long)nib4 << (16 * e);
        }
    }
    const int any = (bv | bh) != 0ull;
    P.bs[(size_t)nmb * 64 + a] = any;
    if (!any) return;
    // thresholds and code look-up tables per (plane, edge kind): kind 0 left MB edge, 1 top MB edge, 2 internal edges
    uint32_t out[DBREC_WORDS];
    unsigned long long lut[3][3];
#pragma unroll
    for (int pl = 0; pl < 3; pl++) {
        uint32_t ab[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int qq = I.qp, qp = k == 0 ? (A >= 0 ? IA.qp : qq) : k == 1 ? (B >= 0 ? IB.qp : qq) : qq;
            if (pl) { qq = chroma_qp(P, qq, pl - 1); qp = chroma_qp(P, qp, pl - 1); }
            const int qpav = (qp + qq + 1) >> 1;
            const uint32_t ta = s_ta[clip3i(0, 51, qpav + I.offa)], tb = s_tb[clip3i(0, 51, qpav + I.offb)];
            ab[k] = (ta & 0xffu) | (tb << 8);
            lut[pl][k] = ((unsigned long long)(((ta >> 8) & 31u) | 0x40u) << 8) | ((unsigned long long)(((ta >> 13) & 31u) | 0x40u) << 16) |
                         ((unsigned long long)(((ta >> 18) & 31u) | 0x40u) << 24) | (0xC0ull << 32);
        }
        out[16 + 2 * pl] = ab[0] | (ab[1] << 16);
        out[17 + 2 * pl] = ab[2];
    }
    out[22] = out[23] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {                                          // luma roles
        uint32_t wv = 0, wh = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            wv |= bs_code(lut[0][s == 0 ? 0 : 2], (uint32_t)(bv >> (4 * (s * 4 + i))) & 15u) << (8 * s);
            wh |= bs_code(lut[0][s == 0 ? 1 : 2], (uint32_t)(bh >> (4 * (s * 4 + i))) & 15u) << (8 * s);
        }
        out[2 * i] = wv; out[2 * i + 1] = wh;
    }
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int h = 0; h < 2; h++) {                                      // chroma roles: half h covers luma segments 2h, 2h + 1
            uint32_t wv = 0, wh = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int e = (t >> 1) * 2, seg = 2 * h + (t & 1);         // chroma edge 4 = luma edge 8 (DB:871-880)
                wv |= bs_code(lut[1 + c][e == 0 ? 0 : 2], (uint32_t)(bv >> (4 * (e * 4 + seg))) & 15u) << (8 * t);
                wh |= bs_code(lut[1 + c][e == 0 ? 1 : 2], (uint32_t)(bh >> (4 * (e * 4 + seg))) & 15u) << (8 * t);
            }
            out[2 * (4 + 2 * c + h)] = wv; out[2 * (4 + 2 * c + h) + 1] = wh;
        }
    uint4 *rec = (uint4 *)(P.bs + (size_t)a * DBREC_WORDS);
#pragma unroll
    for (int i = 0; i < DBREC_WORDS / 4; i++) rec[i] = make_uint4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
}

// ------------------------------------------------------------------ the filter kernel
#define DB3_ROWS 8                       // MB rows (warps) per CTA
#define DB3_THREADS (DB3_ROWS * 32)
#define DB3_SLOTS 43                     // 4x4-sample blocks of one macroblock with its neighbours: 5x5 luma + 2 x 3x3 chroma

// Shared tile of one macroblock (per picture, double-buffered so that macroblock x+1 is staged while x is filtered).  A slot
// is one 4x4-sample block = four row words.  Luma block (br, bc), br/bc = -1 (rows above / columns left) .. 3: slot
// (br+1)*5 + (bc+1); plane c chroma block (br, bc), -1 .. 1: slot 25 + 9c + (br+1)*3 + (bc+1).  Vertical-edge lanes own a
// block ROW (they walk bc), horizontal-edge lanes a block COLUMN (they walk br): the hand-over between the phases is the
// tile itself, no copy.
struct DbTile3 {
    uint4 blk[2][DB_PPW][DB3_SLOTS];
    uint32_t work[8];                 // bit x: some picture of the bundle has a non-zero strength in macroblock x of this row
    uint32_t work_pic[DB_PPW][8];
};
__device__ __forceinline__ uint32_t ldcg32(const void *p) { return __ldcg((const uint32_t *)p); }

struct DbKind { uint32_t ka, kb, ka4; };
__device__ __forceinline__ DbKind db_kind(uint32_t alpha, uint32_t beta) {
    DbKind k;
    k.ka = 0x80008000u - alpha * S16_K1; k.kb = 0x80008000u - beta * S16_K1; k.ka4 = 0x80008000u - ((alpha >> 2) + 2u) * S16_K1;
    return k;
}

// One phase of one macroblock: step s filters the edge between the lane's blocks s-1 and s (luma) — chroma lanes have two
// blocks and filter in steps 0 and 2, so steps 1 and 3 (the edges a transform_size_8x8 macroblock does not have) vanish for
// the whole warp when no luma lane needs them.  The loop is NOT unrolled: one copy of the filter per phase keeps the kernel
// inside the instruction cache (the unrolled first version stalled on instruction fetch as much as on data, ncu).
template <bool VERT>
__device__ __forceinline__ void db3_phase(uint4 *tile, int base, int mul, bool chroma, uint32_t codes, const DbKind &k0, const DbKind &kI, uint32_t lum) {
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        uint32_t pair = s16_prmt(codes, 0u, (chroma ? 0x4140u : 0x4040u) + (uint32_t)s * 0x0101u);
        if (chroma && (s & 1)) pair = 0u;
        const uint32_t act = s16_sign(pair << 9), s4 = s16_sign(pair << 8);
        const bool do3 = __any_sync(0xffffffffu, (act & ~s4) != 0u), do4 = __any_sync(0xffffffffu, s4 != 0u);
        if (!(do3 || do4)) continue;
        DbPar2 k;
        k.kalpha = s == 0 ? k0.ka : kI.ka; k.kbeta = s == 0 ? k0.kb : kI.kb; k.kalpha4 = s == 0 ? k0.ka4 : kI.ka4;
        k.tc0 = pair & 0x001F001Fu; k.act = act; k.s4 = s4; k.lum = lum;
        const int ps = base + mul * (chroma ? (s >> 1) : s), qs = ps + mul;
        const uint4 pv = tile[ps], qv = tile[qs];
        uint32_t P[4] = { pv.x, pv.y, pv.z, pv.w }, Q[4] = { qv.x, qv.y, qv.z, qv.w };
        uint32_t Pe[4], Po[4], Qe[4], Qo[4];
        if (VERT) { blk_rows_to_colpairs(P, Pe, Po); blk_rows_to_colpairs(Q, Qe, Qo); }
        else { blk_rows_to_rowpairs(P, Pe, Po); blk_rows_to_rowpairs(Q, Qe, Qo); }
        db_filter2(Pe[0], Pe[1], Pe[2], Pe[3], Qe[0], Qe[1], Qe[2], Qe[3], k, do3, do4);
        db_filter2(Po[0], Po[1], Po[2], Po[3], Qo[0], Qo[1], Qo[2], Qo[3], k, do3, do4);
        if (VERT) { blk_colpairs_to_rows(Pe, Po, P); blk_colpairs_to_rows(Qe, Qo, Q); }
        else { blk_rowpairs_to_rows(Pe, Po, P); blk_rowpairs_to_rows(Qe, Qo, Q); }
        tile[ps] = make_uint4(P[0], P[1], P[2], P[3]);
        tile[qs] = make_uint4(Q[0], Q[1], Q[2], Q[3]);
    }
}

__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(saddr), "l"(g) : "memory"); }

#ifdef DB3_CLOCKS
#define DB3_T(i) do { const long long t_ = clock64(); sec[i] += t_ - tlast; tlast = t_; } while (0)
#else
#define DB3_T(i) do { } while (0)
#endif
// grid: CTAs draw (band, bundle) tickets; block: DB3_ROWS warps = consecutive MB rows of one bundle of DB_PPW pictures
#ifndef DB3_MIN_CTAS
#define DB3_MIN_CTAS 3                   // 80 registers: the 64-register build spills inside the filter loop and is 25 % slower (measured)
#endif
__global__ void __launch_bounds__(DB3_THREADS, DB3_MIN_CTAS) k_deblock3(const PicDev *pics, int npics, int bands, int *ticket) {
    __shared__ DbTile3 tiles[DB3_ROWS];
    __shared__ int s_prog[DB3_ROWS];
    __shared__ uint64_t s_bar[DB3_ROWS];
    __shared__ int s_ticket;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    if (threadIdx.x < DB3_ROWS) { s_prog[threadIdx.x] = 0; mbar_init(&s_bar[threadIdx.x], 1); }
    __syncthreads();
    const int nbundles = (npics + DB_PPW - 1) / DB_PPW;
    const int tk = s_ticket;
    if (tk >= nbundles * bands) return;
    const int bundle = tk % nbundles;
    const int row = (tk / nbundles) * DB3_ROWS + warp;
    const PicDev &P0 = pics[bundle * DB_PPW];
    const int wmb = P0.wmb, hmb = P0.hmb, nmb = wmb * hmb;
    if (row >= hmb) return;
    const int W = wmb * 16, H = hmb * 16, Wc = W >> 1;
    DbTile3 &T = tiles[warp];
    RowSync rs = rs_init(s_prog, s_bar, warp, row, hmb, P0.progress + hmb, wmb, DB3_ROWS);      // the bundle advances on its first picture's counters

    // ---- work bitmaps of this row
    for (int p = 0; p < DB_PPW; p++) {
        const int pi = bundle * DB_PPW + p;
        const PicDev &Pp = pics[min(pi, npics - 1)];
        const bool pv = pi < npics && Pp.deblock_enable;
        for (int g = 0; g < 8; g++) {
            const int xl = g * 32 + lane;
            int wk = 0;
            if (pv && xl < wmb) { const int a = row * wmb + xl; if (a < Pp.deblock_stop) wk = Pp.bs[(size_t)nmb * 64 + a] != 0u; }
            const unsigned m = __ballot_sync(0xffffffffu, wk);
            if (lane == 0) { T.work_pic[p][g] = m; T.work[g] = p ? (T.work[g] | m) : m; }
        }
    }
    __syncwarp();

    // ---- lane roles
    const int pic = lane >> 3, sub = lane & 7;
    const bool chroma = sub >= 4;
    const int cpl = (sub - 4) >> 1, hw = sub & 1;                          // chroma: plane, and half (vertical phase: rows 4hw..) / word (horizontal phase)
    const uint32_t lum = chroma ? 0u : 0xFFFFFFFFu;
    const bool valid = bundle * DB_PPW + pic < npics;
    const PicDev &P = pics[valid ? bundle * DB_PPW + pic : bundle * DB_PPW];
    uint8_t *plane = P.dst + (chroma ? (size_t)W * H + (cpl ? (size_t)Wc * (H >> 1) : 0) : 0);
    const int rstride = chroma ? Wc : W, mbw = chroma ? 8 : 16;
    // vertical phase: this lane's four rows; horizontal phase: this lane's 4-column word, starting 4 rows above the macroblock
    uint8_t *vrow = plane + (size_t)(row * mbw + 4 * (chroma ? hw : sub)) * rstride;
    uint8_t *hcol = plane + (size_t)(row * mbw - 4) * rstride + 4 * (chroma ? hw : sub);
    const uint32_t *recs = P.bs + (size_t)row * wmb * DBREC_WORDS + 2 * sub;
    const uint32_t *thrs = P.bs + (size_t)row * wmb * DBREC_WORDS + 16 + 2 * (chroma ? 1 + cpl : 0);
    const int vbase = chroma ? 25 + 9 * cpl + (hw + 1) * 3 : (sub + 1) * 5;     // this lane's block row: slots vbase .. vbase+4 (chroma +2)
    const int hbase = chroma ? 25 + 9 * cpl + hw + 1 : sub + 1;                  // this lane's block column: slots hbase + k*hmul
    const int hmul = chroma ? 3 : 5;

    // Prefetch of macroblock x into tile buffer b: the samples go straight from global to shared memory (cp.async, 4 bytes per
    // (row, word): the tile is block-major), so no register carries them across the filtering of the current macroblock; the two
    // record words travel in registers.  Own rows and the columns left of them are not written by any other warp before this
    // warp filters them, so the L1-allocating form is safe.
    auto fetch = [&](int x, bool need_left, int b, uint2 &codes, uint2 &thr) {
        codes = make_uint2(0u, 0u); thr = make_uint2(0u, 0u);
        if (valid) {
            const uint8_t *r0 = vrow + (size_t)x * mbw;
            const uint32_t dst = smem_addr(&T.blk[b][pic][vbase]);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint8_t *rp = r0 + (size_t)t * rstride;
                cp_async4(dst + 16 + 4 * t, rp); cp_async4(dst + 32 + 4 * t, rp + 4);
                if (!chroma) { cp_async4(dst + 48 + 4 * t, rp + 8); cp_async4(dst + 64 + 4 * t, rp + 12); }
                if (need_left && x > 0) cp_async4(dst + 4 * t, rp - 4);
            }
            const bool wk = (T.work_pic[pic][x >> 5] >> (x & 31)) & 1u;
            if (wk) { codes = *(const uint2 *)(recs + (size_t)x * DBREC_WORDS); thr = *(const uint2 *)(thrs + (size_t)x * DBREC_WORDS); }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

#ifdef DB3_CLOCKS
    long long sec[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64(); int iters = 0;
#endif
    int x = db_next_work(T.work, -1, wmb);
    int buf = 0;
    uint2 codes = make_uint2(0u, 0u), thr = make_uint2(0u, 0u), ncodes = codes, nthr = thr;
    if (x < wmb) fetch(x, true, 0, codes, thr);
    while (x < wmb) {
        uint4 *tile = T.blk[buf][pic];
        const int xn = db_next_work(T.work, x, wmb);
        const bool carry = xn == x + 1;
        DB3_T(7);
        asm volatile("cp.async.wait_group 0;" ::: "memory");              // macroblock x has landed in tile[buf] (this lane's own copies)
        DB3_T(0);
        if (xn < wmb) fetch(xn, !carry, buf ^ 1, ncodes, nthr);
        const int topf = __any_sync(0xffffffffu, (s16_prmt(codes.y, 0u, chroma ? 0x4140u : 0x4040u) & 0x00400040u) != 0u);
        // rows above: fetch them early when the row above is already far enough, so that their latency hides behind the vertical phase
        uint32_t Ab[4] = {0, 0, 0, 0};
        const uint8_t *ap = hcol + (size_t)x * mbw;
        const bool early = topf && rs_try(rs, min(x + 2, wmb), lane);
        if (early && valid) {
#pragma unroll
            for (int t = 0; t < 4; t++) Ab[t] = ldcg32(ap + (size_t)t * rstride);
        }
        const DbKind kI = db_kind(thr.y & 0xffu, (thr.y >> 8) & 0xffu);
        DB3_T(1);
        // ---- vertical edges (lanes own block rows)
        {
            const DbKind kL = db_kind(thr.x & 0xffu, (thr.x >> 8) & 0xffu);
            db3_phase<true>(tile, vbase, 1, chroma, codes.x, kL, kI, lum);
        }
        DB3_T(2);
        if (topf) {
            if (!early) {
                rs_wait(rs, min(x + 2, wmb), x, lane);
                if (valid) {
#pragma unroll
                    for (int t = 0; t < 4; t++) Ab[t] = ldcg32(ap + (size_t)t * rstride);
                }
            }
            tile[hbase] = make_uint4(Ab[0], Ab[1], Ab[2], Ab[3]);
        }
        __syncwarp();
        DB3_T(3);
        // ---- horizontal edges (lanes own block columns)
        {
            const DbKind kT = db_kind((thr.x >> 16) & 0xffu, thr.x >> 24);
            db3_phase<false>(tile, hbase, hmul, chroma, codes.y, kT, kI, lum);
        }
        DB3_T(4);
        if (topf && valid) {
            const uint4 a = tile[hbase];
            uint8_t *o = hcol + (size_t)x * mbw;
            *(uint32_t *)(o + (size_t)rstride) = a.y; *(uint32_t *)(o + (size_t)2 * rstride) = a.z; *(uint32_t *)(o + (size_t)3 * rstride) = a.w;
        }
        __syncwarp();
        // ---- final rows of this macroblock (row layout again), the finished columns left of it, the carry into the next macroblock
        {
            uint8_t *r0 = vrow + (size_t)x * mbw;
            const uint4 b0 = tile[vbase + 1], b1 = tile[vbase + 2];
            uint4 last = b1;
            if (!chroma) {
                const uint4 b2 = tile[vbase + 3], b3 = tile[vbase + 4];
                last = b3;
                if (valid) {
                    *(uint4 *)(r0) = make_uint4(b0.x, b1.x, b2.x, b3.x); *(uint4 *)(r0 + (size_t)rstride) = make_uint4(b0.y, b1.y, b2.y, b3.y);
                    *(uint4 *)(r0 + (size_t)2 * rstride) = make_uint4(b0.z, b1.z, b2.z, b3.z); *(uint4 *)(r0 + (size_t)3 * rstride) = make_uint4(b0.w, b1.w, b2.w, b3.w);
                }
            } else if (valid) {
                *(uint2 *)(r0) = make_uint2(b0.x, b1.x); *(uint2 *)(r0 + (size_t)rstride) = make_uint2(b0.y, b1.y);
                *(uint2 *)(r0 + (size_t)2 * rstride) = make_uint2(b0.z, b1.z); *(uint2 *)(r0 + (size_t)3 * rstride) = make_uint2(b0.w, b1.w);
            }
            if (valid && x > 0) {
                const uint4 l = tile[vbase];
                *(uint32_t *)(r0 - 4) = l.x; *(uint32_t *)(r0 + (size_t)rstride - 4) = l.y; *(uint32_t *)(r0 + (size_t)2 * rstride - 4) = l.z; *(uint32_t *)(r0 + (size_t)3 * rstride - 4) = l.w;
            }
            if (carry) T.blk[buf ^ 1][pic][vbase] = last;
        }
        codes = ncodes; thr = nthr;
        DB3_T(5);
        rs_publish(rs, carry ? x + 1 : min(xn, wmb), lane);      // no work up to xn: those columns are final as they are
        DB3_T(6);
#ifdef DB3_CLOCKS
        iters++;
#endif
        buf ^= 1;
        x = xn;
    }
#ifdef DB3_CLOCKS
    if (lane == 0 && bundle == 0 && (row == 20 || row == 23 || row == 24)) printf("DB3CLK row %d iters %d: cpwait %lld fetch+top %lld V %lld topwait %lld H %lld final %lld publish %lld loop %lld\n", row, iters, sec[0] / iters, sec[1] / iters, sec[2] / iters, sec[3] / iters, sec[4] / iters, sec[5] / iters, sec[6] / iters, sec[7] / iters);
#endif
    rs_publish(rs, wmb, lane);
}
