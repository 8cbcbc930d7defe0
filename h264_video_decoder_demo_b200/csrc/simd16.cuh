// simd16.cuh — the deblocking edge filter on TWO sample lines per instruction (packed 2 x 16 bit), plus the 4x4 byte-block
// re-formatting around it.  Compiles for the device (sm_100a: VABSDIFF4, VIMNMX.U16x2, PRMT, LOP3, IADD3) and, with plain C
// stand-ins for the three intrinsics, for the host — tests/test_simd_filter.py holds it against the scalar filter equations
// on the CPU (DB:1373-1522 of the reference) before it ever runs on a GPU.
//
// Representation: a "pair" is a uint32 whose two 16-bit halves hold one 8-bit sample each (high bytes zero), the same sample
// position (p3 .. q3) of two different sample lines.  All intermediate sums are biased so that they stay non-negative and
// below 2^15 per half: ordinary 32-bit adds never carry from the low half into the high half.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define S16_FN __device__ __forceinline__
S16_FN uint32_t s16_prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }
S16_FN uint32_t s16_absdiff(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
S16_FN uint32_t s16_minu(uint32_t a, uint32_t b) { return __vminu2(a, b); }
S16_FN uint32_t s16_maxu(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
#else
#define S16_FN static inline
S16_FN uint32_t s16_prmt(uint32_t a, uint32_t b, uint32_t sel) {           // prmt.b32, generic mode (nibble bit 3: replicate the sign)
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t n = (sel >> (4 * i)) & 15u;
        uint32_t byte = (uint32_t)(src >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        d |= byte << (8 * i);
    }
    return d;
}
S16_FN uint32_t s16_absdiff(uint32_t a, uint32_t b) {
    uint32_t d = 0;
    for (int i = 0; i < 4; i++) { const int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff; d |= (uint32_t)(x > y ? x - y : y - x) << (8 * i); }
    return d;
}
S16_FN uint32_t s16_minu(uint32_t a, uint32_t b) { const uint32_t l = (a & 0xffffu) < (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu), h = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16); return l | (h << 16); }
S16_FN uint32_t s16_maxu(uint32_t a, uint32_t b) { const uint32_t l = (a & 0xffffu) > (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu), h = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16); return l | (h << 16); }
#endif

#define S16_K1   0x00010001u
#define S16_KFF  0x00FF00FFu
#define S16_K256 0x01000100u

S16_FN uint32_t s16_sign(uint32_t x) { return s16_prmt(x, 0u, 0xbb99u); }              // 0xFFFF in every half whose bit 15 is set
S16_FN uint32_t s16_sel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }
// threshold constant for "x >= thr" tests on pairs of 8-bit values: bit 15 of (x + s16_ge_k(thr)) is set iff x >= thr
S16_FN uint32_t s16_ge_k(uint32_t thr) { return (0x8000u - thr) * S16_K1; }

// Parameters of one filter call (two sample lines).  kalpha/kbeta/kalpha4: s16_ge_k(alpha), s16_ge_k(beta), s16_ge_k((alpha >> 2) + 2).
// tc0: per half, tC0 of the half's boundary strength (bS 1..3).  act / s4: 0xFFFF in halves with bS > 0 / bS == 4.
// lum: all ones for luma lines, 0 for chroma lines (chromaStyleFilteringFlag, DB:959).
struct DbPar2 { uint32_t kalpha, kbeta, kalpha4, tc0, act, s4, lum; };

// DB:1373-1478 (bS < 4) and DB:1481-1522 (bS == 4) for two sample lines at once.  No data-dependent branch: do3 / do4 say whether
// any line of the CALLER'S WARP has 0 < bS < 4 / bS == 4 (warp-uniform, from the step codes alone), so the only branches are uniform.
S16_FN void db_filter2(uint32_t &P3, uint32_t &P2, uint32_t &P1, uint32_t &P0, uint32_t &Q0, uint32_t &Q1, uint32_t &Q2, uint32_t &Q3, const DbPar2 &k, bool do3, bool do4) {
    const uint32_t dpq = s16_absdiff(P0, Q0);
    const uint32_t ge_a = s16_sign(dpq + k.kalpha);
    const uint32_t ge_b = s16_sign(s16_maxu(s16_absdiff(P1, P0), s16_absdiff(Q1, Q0)) + k.kbeta);
    const uint32_t cond = k.act & ~(ge_a | ge_b);                                      // filterSamplesFlag (DB:1366) of lines with bS > 0
    const uint32_t apm = ~s16_sign(s16_absdiff(P2, P0) + k.kbeta) & k.lum;             // ap < beta, luma only
    const uint32_t aqm = ~s16_sign(s16_absdiff(Q2, Q0) + k.kbeta) & k.lum;
    const uint32_t m3 = cond & ~k.s4, m4 = cond & k.s4;
    uint32_t nP0 = P0, nP1 = P1, nP2 = P2, nQ0 = Q0, nQ1 = Q1, nQ2 = Q2;
    if (do3) {
        const uint32_t tc = k.tc0 + (apm & S16_K1) + (aqm & S16_K1) + (~k.lum & S16_K1);
        // ((q0 - p0) * 4 + (p1 - q1) + 4) + 2048 with every term non-negative; >> 3 gives delta + 256
        const uint32_t t = ((Q0 + (P0 ^ S16_KFF)) << 2) + P1 + (Q1 ^ S16_KFF) + 777u * S16_K1;
        uint32_t d = (t >> 3) & 0x1FFF1FFFu;
        d = s16_minu(s16_maxu(d, S16_K256 - tc), S16_K256 + tc);
        const uint32_t a0 = s16_minu(s16_maxu(P0 + d, S16_K256), 0x01FF01FFu) & S16_KFF;                  // Clip1(p0 + delta)
        const uint32_t b0 = s16_minu(s16_maxu(Q0 + (0x02000200u - d), S16_K256), 0x01FF01FFu) & S16_KFF;    // Clip1(q0 - delta)
        const uint32_t avg = ((P0 + Q0 + S16_K1) >> 1) & S16_KFF;
        const uint32_t lo0 = S16_K256 - k.tc0, hi0 = S16_K256 + k.tc0;
        uint32_t u = ((P2 + avg + ((P1 ^ S16_KFF) << 1) + 2u * S16_K1) >> 1) & 0x01FF01FFu;                // ((p2 + avg - 2 p1) >> 1) + 256
        u = s16_minu(s16_maxu(u, lo0), hi0);
        const uint32_t a1 = P1 + u - S16_K256;
        uint32_t v = ((Q2 + avg + ((Q1 ^ S16_KFF) << 1) + 2u * S16_K1) >> 1) & 0x01FF01FFu;
        v = s16_minu(s16_maxu(v, lo0), hi0);
        const uint32_t b1 = Q1 + v - S16_K256;
        nP0 = s16_sel(m3, a0, nP0); nQ0 = s16_sel(m3, b0, nQ0);
        nP1 = s16_sel(m3 & apm, a1, nP1); nQ1 = s16_sel(m3 & aqm, b1, nQ1);
    }
    if (do4) {
        const uint32_t small = ~s16_sign(dpq + k.kalpha4);                             // |p0 - q0| < (alpha >> 2) + 2
        const uint32_t ps = m4 & apm & small, qs = m4 & aqm & small;
        const uint32_t S = P0 + Q0, Tp = P1 + S, Tq = Q1 + S;
        const uint32_t p0s = ((P2 + 2u * Tp + Q1 + 4u * S16_K1) >> 3) & S16_KFF;
        const uint32_t p1s = ((P2 + Tp + 2u * S16_K1) >> 2) & S16_KFF;
        const uint32_t p2s = ((2u * P3 + 3u * P2 + Tp + 4u * S16_K1) >> 3) & S16_KFF;
        const uint32_t p0w = ((2u * P1 + P0 + Q1 + 2u * S16_K1) >> 2) & S16_KFF;
        const uint32_t q0s = ((Q2 + 2u * Tq + P1 + 4u * S16_K1) >> 3) & S16_KFF;
        const uint32_t q1s = ((Q2 + Tq + 2u * S16_K1) >> 2) & S16_KFF;
        const uint32_t q2s = ((2u * Q3 + 3u * Q2 + Tq + 4u * S16_K1) >> 3) & S16_KFF;
        const uint32_t q0w = ((2u * Q1 + Q0 + P1 + 2u * S16_K1) >> 2) & S16_KFF;
        nP0 = s16_sel(m4, s16_sel(ps, p0s, p0w), nP0); nP1 = s16_sel(ps, p1s, nP1); nP2 = s16_sel(ps, p2s, nP2);
        nQ0 = s16_sel(m4, s16_sel(qs, q0s, q0w), nQ0); nQ1 = s16_sel(qs, q1s, nQ1); nQ2 = s16_sel(qs, q2s, nQ2);
    }
    P0 = nP0; P1 = nP1; P2 = nP2; Q0 = nQ0; Q1 = nQ1; Q2 = nQ2;
}

// ---- 4x4 byte blocks.  A block arrives as four row words r[0..3] (byte x of r[t] = sample (row t, column x)).
// "Pairs" of a block: e[c] = column c of rows (0, 2), o[c] = column c of rows (1, 3) — what the vertical-edge filter wants
// (the two lines of a call are rows, the sample positions are columns) ...
S16_FN void blk_rows_to_colpairs(const uint32_t r[4], uint32_t e[4], uint32_t o[4]) {
    const uint32_t a0 = s16_prmt(r[0], 0u, 0x4240u), b0 = s16_prmt(r[0], 0u, 0x4341u);     // row 0: columns (0, 2) / (1, 3)
    const uint32_t a1 = s16_prmt(r[1], 0u, 0x4240u), b1 = s16_prmt(r[1], 0u, 0x4341u);
    const uint32_t a2 = s16_prmt(r[2], 0u, 0x4240u), b2 = s16_prmt(r[2], 0u, 0x4341u);
    const uint32_t a3 = s16_prmt(r[3], 0u, 0x4240u), b3 = s16_prmt(r[3], 0u, 0x4341u);
    e[0] = s16_prmt(a0, a2, 0x5410u); e[2] = s16_prmt(a0, a2, 0x7632u); e[1] = s16_prmt(b0, b2, 0x5410u); e[3] = s16_prmt(b0, b2, 0x7632u);
    o[0] = s16_prmt(a1, a3, 0x5410u); o[2] = s16_prmt(a1, a3, 0x7632u); o[1] = s16_prmt(b1, b3, 0x5410u); o[3] = s16_prmt(b1, b3, 0x7632u);
}
S16_FN void blk_colpairs_to_rows(const uint32_t e[4], const uint32_t o[4], uint32_t r[4]) {
    const uint32_t a0 = s16_prmt(e[0], e[2], 0x5410u), a2 = s16_prmt(e[0], e[2], 0x7632u), b0 = s16_prmt(e[1], e[3], 0x5410u), b2 = s16_prmt(e[1], e[3], 0x7632u);
    const uint32_t a1 = s16_prmt(o[0], o[2], 0x5410u), a3 = s16_prmt(o[0], o[2], 0x7632u), b1 = s16_prmt(o[1], o[3], 0x5410u), b3 = s16_prmt(o[1], o[3], 0x7632u);
    r[0] = s16_prmt(a0, b0, 0x6240u); r[1] = s16_prmt(a1, b1, 0x6240u); r[2] = s16_prmt(a2, b2, 0x6240u); r[3] = s16_prmt(a3, b3, 0x6240u);
}
// ... and e[t] = columns (0, 2) of row t, o[t] = columns (1, 3) of row t — what the horizontal-edge filter wants (the two
// lines of a call are columns, the sample positions are rows).
S16_FN void blk_rows_to_rowpairs(const uint32_t r[4], uint32_t e[4], uint32_t o[4]) {
#pragma unroll
    for (int t = 0; t < 4; t++) { e[t] = s16_prmt(r[t], 0u, 0x4240u); o[t] = s16_prmt(r[t], 0u, 0x4341u); }
}
S16_FN void blk_rowpairs_to_rows(const uint32_t e[4], const uint32_t o[4], uint32_t r[4]) {
#pragma unroll
    for (int t = 0; t < 4; t++) r[t] = s16_prmt(e[t], o[t], 0x6240u);
}
