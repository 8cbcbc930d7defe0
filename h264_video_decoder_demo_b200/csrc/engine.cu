// engine.cu — per-GPU context, launch sequencing and the C ABI of include/h264_recon_b200.h.
//
// One context = one GPU.  A submit reconstructs one picture for each listed stream (pictures of one
// stream are serial in decoding order; different streams are independent and share every launch):
//
//   launch stream
//   [memset of surfaces that must start from zero (PB:53-69)]
//   k_inter_tma   inter prediction + inter residual, reference windows staged by TMA   (warp per 5 consecutive macroblocks)
//   k_inter_list  the inter macroblocks k_inter_tma left on the work list              (warp per macroblock, clamped loads)
//   k_inter       MBAFF pictures / H264B2_INTER_V1                                     (warp per macroblock)
//   k_intra       intra prediction + intra residual: luma and chroma wavefronts        (MB wavefront, one warp per MB row, band CTAs)
//   k_deblock3    in-loop filter, in place, four pictures per warp                     (MB wavefront; k_deblock<true> for MBAFF)
//   look-ahead stream, one batch ahead
//   k_prologue [k_expand, k_unmotion]  descriptors, packed levels / motion records -> plain arrays
//   k_residual    dequantisation + inverse transforms into the residual tiles          (warp per macroblock)
//   k_bs_prog2    boundary strengths as step codes                                      (thread per macroblock; k_bs for MBAFF)
//
// The decoded picture buffer (n_streams x surfaces_per_stream I420 surfaces, Y|Cb|Cr contiguous like
// PB:167-179) lives in HBM for the life of the context.  There is no CPU fallback anywhere.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include <new>

#include "h264_recon_b200.h"
#include "common.cuh"
#include "residual.cuh"
#include "residual_kernel.cuh"
#include "inter_quad.cuh"
#include "inter_tma.cuh"
#include <cudaTypedefs.h>
#include "intra.cuh"
#include "deblock.cuh"
#include "deblock_simd.cuh"

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(-10, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// ------------------------------------------------------------------ checksum kernel
__global__ void k_checksum(const uint8_t *const *surf, size_t nwords, unsigned long long *out) {
    const uint32_t *w = (const uint32_t *)surf[blockIdx.y];
    const unsigned long long K = 0x9E3779B97F4A7C15ULL;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
        acc += ((unsigned long long)w[i] + 1ULL) * ((2ULL * i + 1ULL) * K);
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&out[blockIdx.y], acc);
}

// Batch prologue: pull the picture descriptors out of mapped pinned host memory and zero the wavefront
// counters.  A kernel, not cudaMemcpyAsync/cudaMemsetAsync: copy-engine work queued on the launch stream
// would wait behind the bulk H2D/D2H transfers of neighbouring batches (measured: +4 ms per batch).
__global__ void k_prologue(const uint4 *src, uint4 *dst, int n16, int *progress, int nprog) {
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < nprog; i += blockDim.x) progress[i] = 0;
}
__global__ void k_zero_ints(int *p, int n) { for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0; }
// snapshot n surfaces into the read-back staging buffer (same reason: keep the copy engines for PCIe)
__global__ void k_snapshot(const uint8_t *const *src, uint8_t *dst, size_t n16) {
    const uint4 *s = (const uint4 *)src[blockIdx.y];
    uint4 *d = (uint4 *)dst + (size_t)blockIdx.y * n16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}

// pull host arrays into the device arena with SM loads over PCIe (optional submit path, H264B2_H2D_ZEROCOPY=<ctas>): the
// copy engines then only carry the read-back.  src and dst share their offset modulo 16.
struct PullDesc { const uint8_t *src; uint8_t *dst; size_t bytes; };
__global__ void __launch_bounds__(256) k_pull(const PullDesc *list, int n) {
    for (int e = 0; e < n; e++) {
        const PullDesc d = list[e];
        const size_t head = (16 - ((uintptr_t)d.src & 15)) & 15;
        const size_t h = head < d.bytes ? head : d.bytes;
        const size_t n16 = (d.bytes - h) / 16, tail = d.bytes - h - n16 * 16;
        if (blockIdx.x == 0 && threadIdx.x < h) d.dst[threadIdx.x] = __ldcv(d.src + threadIdx.x);
        if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x - 32 < tail) d.dst[h + n16 * 16 + threadIdx.x - 32] = __ldcv(d.src + h + n16 * 16 + threadIdx.x - 32);
        const uint4 *s = (const uint4 *)(d.src + h);
        uint4 *o = (uint4 *)(d.dst + h);
        const size_t stride = (size_t)gridDim.x * blockDim.x;
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n16; i += 4 * stride) {
            const uint4 a = __ldcv(s + i), b = __ldcv(s + i + stride), c2 = __ldcv(s + i + 2 * stride), d2 = __ldcv(s + i + 3 * stride);
            o[i] = a; o[i + stride] = b; o[i + 2 * stride] = c2; o[i + 3 * stride] = d2;
        }
        for (; i < n16; i += stride) o[i] = __ldcv(s + i);
    }
}

// drain the staging buffer into page-locked HOST memory with SM stores over PCIe (optional read-back path: no copy-engine
// descriptor per picture; a few persistent CTAs are enough, the stores are posted)
__global__ void __launch_bounds__(256) k_drain(const uint8_t *stage, uint8_t *const *dst, size_t n16, int n) {
    for (int p = 0; p < n; p++) {
        const uint4 *s = (const uint4 *)stage + (size_t)p * n16;
        uint4 *d = (uint4 *)dst[p];
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) d[i] = __ldcs(s + i);
    }
}

// ---- packed coefficient transport (include/h264_recon_b200.h: h264b2_pack_coefs).  Entropy-decoded residuals are sparse
// (5.7 % non-zero in the bundled 1080p streams), so the host may send, instead of the dense int16 array, one 16-bit
// significance map per 16-coefficient chunk plus the non-zero levels; k_expand rebuilds the dense array in HBM before
// k_residual runs.  Blob: uint32 {magic, n_chunks, n_values, total_bytes}, uint32 base[ceil(n_chunks/32)] (first value
// index of each group of 32 chunks), uint16 map[n_chunks rounded up to even], int16 values[n_values].
#define H264B2_PACK_MAGIC 0x4B503248u
__global__ void __launch_bounds__(256) k_expand(const PicDev *pics) {
    const PicDev &P = pics[blockIdx.y];
    const uint32_t *blob = blockIdx.z ? P.packed_motion : P.packed;         // grid.z: 0 = coefficients, 1 = motion records
    if (!blob) return;
    int16_t *out = blockIdx.z ? (int16_t *)P.motion : (int16_t *)P.coefs;
    const uint32_t n_chunks = blob[1], n_groups = (n_chunks + 31) >> 5;
    const uint32_t *base = blob + 4;
    const uint16_t *map = (const uint16_t *)(base + n_groups);
    const int16_t *values = (const int16_t *)(map + ((n_chunks + 1) & ~1u));
    const int lane = threadIdx.x & 31;
    for (uint32_t g = blockIdx.x * 8 + (threadIdx.x >> 5); g < n_groups; g += gridDim.x * 8) {
        const uint32_t chunk = g * 32 + lane;
        const uint32_t bm = chunk < n_chunks ? map[chunk] : 0u;
        const int pc = __popc(bm);
        int incl = pc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        const int16_t *v = values + base[g] + (incl - pc);
        if (chunk < n_chunks) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t lo = (bm >> (2 * i)) & 1u ? (uint16_t)v[__popc(bm & ((1u << (2 * i)) - 1))] : 0u;
                const uint32_t hi = (bm >> (2 * i + 1)) & 1u ? (uint16_t)v[__popc(bm & ((1u << (2 * i + 1)) - 1))] : 0u;
                w[i] = lo | (hi << 16);
            }
            uint4 *dst = (uint4 *)(out + (size_t)chunk * 16);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]); dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

// undo the XOR chain h264b2_pack_motion put on the 16 vectors of each list (thread = one list of one macroblock)
__global__ void __launch_bounds__(256) k_unmotion(const PicDev *pics) {
    const PicDev &P = pics[blockIdx.y];
    if (!P.packed_motion) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.wmb * P.hmb * 2) return;
    uint2 *w = (uint2 *)((uint8_t *)P.motion + (size_t)(t >> 1) * sizeof(H264B2MbMotion) + (t & 1) * 64);     // records are 8-byte aligned
    uint32_t prev = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { uint2 v = w[i]; v.x ^= prev; v.y ^= v.x; prev = v.y; w[i] = v; }
}

// YUV420P -> BGR24 (the reference's integer BT.601 conversion, H264PictureBase.cpp:440-468 / FlipLines :471-498).
// One thread = 4 horizontally adjacent pixels: one 32-bit luma load, 12 output bytes as three 32-bit stores.
__global__ void k_bgr24(const uint8_t *i420, int W, int H, uint8_t *bgr, int width_bytes, int flip) {
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x4 * 4 >= W) return;
    const uint32_t yy = *(const uint32_t *)(i420 + (size_t)y * W + x4 * 4);
    const uint8_t *pu = i420 + (size_t)W * H + (size_t)(y >> 1) * (W >> 1) + x4 * 2, *pv = pu + (size_t)W * H / 4;
    const int U0 = pu[0] - 128, U1 = pu[1] - 128, V0 = pv[0] - 128, V1 = pv[1] - 128;
    uint8_t o[12];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int Y = 1164 * ((int)((yy >> (8 * i)) & 0xff) - 16), U = i < 2 ? U0 : U1, V = i < 2 ? V0 : V1;
        o[3 * i + 0] = (uint8_t)clip255((Y + 2018 * U) / 1000);
        o[3 * i + 1] = (uint8_t)clip255((Y - 813 * V - 391 * U) / 1000);
        o[3 * i + 2] = (uint8_t)clip255((Y + 1596 * V) / 1000);
    }
    uint8_t *dst = bgr + (size_t)(flip ? H - 1 - y : y) * width_bytes + x4 * 12;
    if ((((uintptr_t)dst) & 3) == 0) {
        uint32_t *d32 = (uint32_t *)dst;
        d32[0] = o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24);
        d32[1] = o[4] | (o[5] << 8) | (o[6] << 16) | ((uint32_t)o[7] << 24);
        d32[2] = o[8] | (o[9] << 8) | (o[10] << 16) | ((uint32_t)o[11] << 24);
    } else {
#pragma unroll
        for (int i = 0; i < 12; i++) dst[i] = o[i];
    }
}

// ------------------------------------------------------------------ context
enum { NSLOT = H264B2_SUBMIT_DEPTH, NOUT = 4, DESC_RING = 8, EV_POOL = 1 << 17, KCLASSES = 6, MAX_GROUPS = 8 };

struct H264B2Context {
    int device, n_streams, spp, wmb, hmb, nmb;
    size_t frame_bytes;
    uint8_t *surfaces;
    uint32_t *bs;
    int16_t *res;
    int *progress;            // [n_streams][3][hmb] (intra / luma intra, deblock, chroma intra) then DESC_RING * 6 * MAX_GROUPS tickets
    size_t progress_ints;
    int16_t *ls_flat;         // ls4 [2][2][6][16] then ls8 [2][2][6][64]
    cudaStream_t st, st_h2d, st_h2d2, st_d2h;
    int h2d_streams;                          // 1 or 2 copy streams for host-array submits (H264B2_H2D_STREAMS)
    int d2h_zerocopy, d2h_ctas;               // H264B2_D2H_ZEROCOPY=<ctas>: read-backs leave through k_drain instead of the copy engine
    size_t h2d_last_bytes; unsigned h2d_last_copies;     // shape of the last host-array submit (drives the read-back policy)
    size_t d2h_chunk;                         // read-backs into adjacent host memory are merged into copies of at most this size
    // A batch is cut into `groups` picture groups whose kernel sequences run on separate streams, forked from and
    // joined back into `st`: the latency-bound wavefront kernels of one group overlap the issue-bound kernels
    // (and the wavefronts) of the others.
    int groups, group_min;
    cudaStream_t st_g[MAX_GROUPS], st_side[MAX_GROUPS];
    cudaEvent_t fork_ev, join_ev[MAX_GROUPS], side_fork[MAX_GROUPS], side_join[MAX_GROUPS];
    // descriptor ring
    PullDesc *h_pull, *h_pull_dev; int h2d_ctas;     // H264B2_H2D_ZEROCOPY: per-slot lists of host spans for k_pull
    PicDev *h_desc, *h_desc_dev, *d_desc;     // h_desc: mapped pinned host memory, h_desc_dev: its device alias
    cudaEvent_t desc_ev[DESC_RING];
    int desc_next;
    // host-submit staging: device arena slots
    uint8_t *arena[NSLOT]; size_t arena_cap[NSLOT];
    cudaEvent_t h2d_done[NSLOT], h2d_done2[NSLOT], compute_done[NSLOT];
    int slot_next;
    // read-back staging
    uint8_t *out_stage[NOUT]; size_t out_cap[NOUT]; cudaEvent_t out_ready[NOUT], out_done[NOUT]; int out_next;
    uint8_t **d_ptrs; unsigned long long *d_sums; unsigned long long *h_sums; uint8_t **h_ptrs;
    uint8_t **h_snap[NOUT], **h_snap_dev[NOUT];     // mapped pinned pointer lists for k_snapshot
    uint8_t *bgr; size_t bgr_cap;             // BGR24 output staging
    // look-ahead (H264B2_LOOKAHEAD, default on): descriptors, unpacking, k_residual and k_bs of batch i+1 need nothing from batch i,
    // so they run on st_pre while the dependency-bound wavefront kernels of batch i leave issue slots idle; res / bs are double-buffered
    int *worklist; size_t worklist_stride;    // per stream: entry count + addresses of the macroblocks k_inter_tma leaves to k_inter_list
    int inter_tma; CUtensorMap map_y, map_c;   // TMA descriptors of the DPB (luma: x, y, surface; chroma: x, y, plane, surface); H264B2_INTER_V1=1 keeps the LDG kernel
    int debug_launch;                         // H264B2_DEBUG_LAUNCH=1: check for a launch error after every kernel of a batch
    int intra_split; cudaStream_t st_chroma[MAX_GROUPS]; cudaEvent_t chroma_fork[MAX_GROUPS], chroma_join[MAX_GROUPS];   // H264B2_INTRA_SPLIT=0: one intra wavefront for luma and chroma
    int deblock_v1;                           // H264B2_DEBLOCK_V1=1: first-generation progressive deblocking kernels (A/B runs)
    int lookahead; unsigned batch_no; cudaStream_t st_pre; cudaEvent_t pre_done[DESC_RING], main_done[2];
    size_t bs_stride;                         // words per stream in bs
    // timing
    cudaEvent_t t0, t1;
    int timing;
    cudaEvent_t *ev; int ev_used; int ev_class[EV_POOL / 2];
    float class_ms[KCLASSES]; long long class_launches[KCLASSES];
    // optional timeline trace (H264B2_TRACE=path): per host submit, events at H2D start/end and compute start/end
    int trace; int trace_n; cudaEvent_t trace_ev[256][6]; const char *trace_path;
};

static void trace_mark(H264B2Context *c, int which, cudaStream_t s, int back = 0) {
    if (!c->trace || c->trace_n - back >= 256 || c->trace_n - back < 0) return;
    cudaEvent_t &e = c->trace_ev[c->trace_n - back][which];
    if (!e) cudaEventCreate(&e);
    cudaEventRecord(e, s);
}
static void trace_dump(H264B2Context *c) {
    if (!c->trace || c->trace_n < 2) return;
    FILE *f = fopen(c->trace_path, "a");
    if (!f) return;
    fprintf(f, "# batch h2d_start h2d_end compute_start compute_end d2h_start d2h_end (ms since first H2D start)\n");
    for (int i = 0; i < c->trace_n; i++) {
        float t[6] = {0, 0, 0, 0, -1, -1};
        for (int k = 0; k < 6; k++) if (c->trace_ev[i][k]) { if (cudaEventElapsedTime(&t[k], c->trace_ev[0][0], c->trace_ev[i][k]) != cudaSuccess) { cudaGetLastError(); t[k] = -1; } }
        fprintf(f, "%d %.3f %.3f %.3f %.3f %.3f %.3f\n", i, t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    fclose(f);
    c->trace_n = 0;
}

// Directional intra modes 3..8 as (index, index, index, kind) tables (layout: pred_tab_px in intra.cuh).  N = 4: neighbours
// nb[0] corner, nb[1..4] left, nb[5..12] top (PB:1174-1395); N = 8: qa[0] corner, qa[1..8] left, qa[9..24] top (PB:1606-1831).
static void fill_intra_tables(int N, uint16_t *tab) {
    const int LB = 1, TB = 1 + N;                      // first left / first top sample
    auto T = [&](int i) { return i < 0 ? 0 : TB + i; };
    auto L = [&](int i) { return i < 0 ? 0 : LB + i; };
    auto e3 = [](int a, int b, int c) { return (uint16_t)(a | (b << 5) | (c << 10)); };
    auto e2 = [](int a, int b) { return (uint16_t)(a | (b << 5) | (1 << 15)); };
    auto e1 = [](int a) { return (uint16_t)(a | (a << 5) | (a << 10)); };      // copy: (a + 2a + a + 2) >> 2
    for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) {
        uint16_t *o = tab + y * N + x;
        const int NN = N * N;
        // 3: diagonal down left
        o[0 * NN] = (x == N - 1 && y == N - 1) ? e3(T(2 * N - 2), T(2 * N - 1), T(2 * N - 1)) : e3(T(x + y), T(x + y + 1), T(x + y + 2));
        // 4: diagonal down right
        o[1 * NN] = x > y ? e3(T(x - y - 2), T(x - y - 1), T(x - y)) : x < y ? e3(L(y - x - 2), L(y - x - 1), L(y - x)) : e3(T(0), 0, L(0));
        // 5: vertical right
        { const int z = 2 * x - y, k = x - (y >> 1);
          o[2 * NN] = (z >= 0 && !(z & 1)) ? e2(T(k - 1), T(k)) : z >= 0 ? e3(T(k - 2), T(k - 1), T(k)) : z == -1 ? e3(L(0), 0, T(0))
                    : N == 4 ? e3(L(y - 1), L(y - 2), L(y - 3)) : e3(L(y - 2 * x - 1), L(y - 2 * x - 2), L(y - 2 * x - 3)); }
        // 6: horizontal down
        { const int z = 2 * y - x, k = y - (x >> 1);
          o[3 * NN] = (z >= 0 && !(z & 1)) ? e2(L(k - 1), L(k)) : z >= 0 ? e3(L(k - 2), L(k - 1), L(k)) : z == -1 ? e3(L(0), 0, T(0))
                    : N == 4 ? e3(T(x - 1), T(x - 2), T(x - 3)) : e3(T(x - 2 * y - 1), T(x - 2 * y - 2), T(x - 2 * y - 3)); }
        // 7: vertical left
        { const int k = x + (y >> 1); o[4 * NN] = !(y & 1) ? e2(T(k), T(k + 1)) : e3(T(k), T(k + 1), T(k + 2)); }
        // 8: horizontal up
        { const int z = x + 2 * y, k = y + (x >> 1), lim = 2 * N - 3;      // 5 for 4x4, 13 for 8x8
          o[5 * NN] = (z < lim && !(z & 1)) ? e2(L(k), L(k + 1)) : z < lim ? e3(L(k), L(k + 1), L(k + 2)) : z == lim ? e3(L(N - 2), L(N - 1), L(N - 1)) : e1(L(N - 1)); }
    }
}

static const uint8_t h_zz4[16] = {0,1,4,8, 5,2,3,6, 9,12,13,10, 7,11,14,15};
static const uint8_t h_fs4[16] = {0,4,1,8, 12,5,9,13, 2,6,10,14, 3,7,11,15};
static const uint8_t h_fs8[64] = {
     0, 8,16, 1, 9,24,32,17,  2,25,40,48,56,33,10, 3,
    18,41,49,57,26,11, 4,19, 34,42,50,58,27,12, 5,20,
    35,43,51,59,28,13, 6,21, 36,44,52,60,29,14,22,37,
    45,53,61,30, 7,15,38,46, 54,62,23,31,39,47,55,63 };
static const int h_na4[6][3] = {{10,16,13},{11,18,14},{13,20,16},{14,23,18},{16,25,20},{18,29,23}};
static const int h_na8[6][6] = {{20,18,32,19,25,24},{22,19,35,21,28,26},{26,23,42,24,33,31},{28,25,45,26,35,33},{32,28,51,30,40,38},{36,32,58,34,46,43}};
static int na4(int m, int pos) { int i = pos >> 2, j = pos & 3; return (!(i & 1) && !(j & 1)) ? h_na4[m][0] : ((i & 1) && (j & 1)) ? h_na4[m][1] : h_na4[m][2]; }
static int na8(int m, int pos) {
    int i = pos >> 3, j = pos & 7;
    if (i % 4 == 0 && j % 4 == 0) return h_na8[m][0];
    if (i % 2 == 1 && j % 2 == 1) return h_na8[m][1];
    if (i % 4 == 2 && j % 4 == 2) return h_na8[m][2];
    if ((i % 4 == 0 && j % 2 == 1) || (i % 2 == 1 && j % 4 == 0)) return h_na8[m][3];
    if ((i % 4 == 0 && j % 4 == 2) || (i % 4 == 2 && j % 4 == 0)) return h_na8[m][4];
    return h_na8[m][5];
}

static int init_tables(H264B2Context *c) {
    uint8_t zz8[64];
    { int i = 0, j = 0;
      for (int k = 0; k < 64; k++) {
          zz8[k] = (uint8_t)(i * 8 + j);
          if ((i + j) % 2 == 0) { if (j == 7) i++; else if (i == 0) j++; else { i--; j++; } }
          else { if (i == 7) j++; else if (j == 0) i++; else { i++; j--; } }
      } }
    uint8_t is4[2][16], is8[2][64];
    for (int k = 0; k < 16; k++) { is4[0][h_zz4[k]] = (uint8_t)k; is4[1][h_fs4[k]] = (uint8_t)k; }
    for (int k = 0; k < 64; k++) { is8[0][zz8[k]] = (uint8_t)k; is8[1][h_fs8[k]] = (uint8_t)k; }
    { uint16_t t4[6][16], t8[6][64];
      fill_intra_tables(4, &t4[0][0]); fill_intra_tables(8, &t8[0][0]);
      CK(cudaMemcpyToSymbol(g_pred4_tab, t4, sizeof t4)); CK(cudaMemcpyToSymbol(g_pred8_tab, t8, sizeof t8)); }
    CK(cudaMemcpyToSymbol(c_iscan4, is4, sizeof is4));
    CK(cudaMemcpyToSymbol(c_iscan8, is8, sizeof is8));
    // Flat_4x4_16 / Flat_8x8_16 LevelScale in list order (PB:4852-4989 with weightScale = 16)
    std::vector<int16_t> ls(2 * 2 * 6 * 16 + 2 * 2 * 6 * 64);
    size_t o = 0;
    for (int inter = 0; inter < 2; inter++) for (int f = 0; f < 2; f++) for (int m = 0; m < 6; m++)
        for (int k = 0; k < 16; k++) ls[o++] = (int16_t)(16 * na4(m, f ? h_fs4[k] : h_zz4[k]));
    for (int inter = 0; inter < 2; inter++) for (int f = 0; f < 2; f++) for (int m = 0; m < 6; m++)
        for (int k = 0; k < 64; k++) ls[o++] = (int16_t)(16 * na8(m, f ? h_fs8[k] : zz8[k]));
    CK(cudaMalloc(&c->ls_flat, ls.size() * 2));
    CK(cudaMemcpy(c->ls_flat, ls.data(), ls.size() * 2, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int h264b2_abi_version(void) { return H264B2_ABI_VERSION; }
extern "C" const char *h264b2_last_error(void) { return g_err; }

extern "C" uint64_t h264b2_checksum_host(const uint8_t *data, size_t bytes) {
    const uint64_t K = 0x9E3779B97F4A7C15ULL;
    uint64_t acc = 0; const size_t nw = bytes / 4;
    for (size_t i = 0; i < nw; i++) { uint32_t w; memcpy(&w, data + 4 * i, 4); acc += ((uint64_t)w + 1) * ((2 * (uint64_t)i + 1) * K); }
    return acc;
}

static int create_impl(H264B2Context *c, int device, int n_streams, int surfaces_per_stream, int width_mbs, int height_mbs);
extern "C" int h264b2_create(H264B2Context **out, int device, int n_streams, int surfaces_per_stream, int width_mbs, int height_mbs) {
    if (!out || n_streams < 1 || surfaces_per_stream < 1 || surfaces_per_stream > 32 || width_mbs < 1 || height_mbs < 1)
        return fail(-1, "h264b2_create: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(-11, "h264b2_create: no CUDA device (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(-1, "h264b2_create: device %d out of range", device);
    CK(cudaSetDevice(device));
    H264B2Context *c = new (std::nothrow) H264B2Context();
    if (!c) return fail(-12, "out of memory");
    memset(c, 0, sizeof *c);
    const int rc = create_impl(c, device, n_streams, surfaces_per_stream, width_mbs, height_mbs);
    if (rc) {                                   // release whatever was created before the failing call (the error text stays)
        char keep[sizeof g_err]; memcpy(keep, g_err, sizeof keep);
        h264b2_destroy(c);
        cudaGetLastError();
        memcpy(g_err, keep, sizeof keep);
        return rc;
    }
    *out = c;
    return 0;
}
static int create_impl(H264B2Context *c, int device, int n_streams, int surfaces_per_stream, int width_mbs, int height_mbs) {
    c->device = device; c->n_streams = n_streams; c->spp = surfaces_per_stream; c->wmb = width_mbs; c->hmb = height_mbs;
    c->nmb = width_mbs * height_mbs; c->frame_bytes = (size_t)c->nmb * 384;
    const size_t total = (size_t)n_streams * surfaces_per_stream * c->frame_bytes;
    // tail padding: a bottom-field view clamps x to 2W-1 and may read one row past the last plane (Q4)
    CK(cudaMalloc(&c->surfaces, total + (size_t)width_mbs * 64));
    CK(cudaMemset(c->surfaces, 0, total + (size_t)width_mbs * 64));
    // per stream: 64 words of strengths per MB + 1 "any strength" word per MB; the stride is rounded to 16 bytes because k_bs
    // writes the records with 16-byte stores (an odd macroblock count would misalign every second stream otherwise)
    c->bs_stride = ((size_t)c->nmb * 65 + 3) & ~(size_t)3;
    { const char *e = getenv("H264B2_LOOKAHEAD"); c->lookahead = !(e && atoi(e) == 0); }
    { const char *e = getenv("H264B2_DEBUG_LAUNCH"); c->debug_launch = e && atoi(e) != 0; }
    { const char *e = getenv("H264B2_DEBLOCK_V1"); c->deblock_v1 = e && atoi(e) != 0; }
    CK(cudaMalloc(&c->bs, (size_t)n_streams * c->bs_stride * 4 * 2));
    CK(cudaMalloc(&c->res, (size_t)n_streams * c->nmb * RES_MB_STRIDE * 2 * 2));
    CK(cudaStreamCreateWithFlags(&c->st_pre, cudaStreamNonBlocking));
    for (int i = 0; i < DESC_RING; i++) CK(cudaEventCreateWithFlags(&c->pre_done[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&c->main_done[i], cudaEventDisableTiming));
    c->progress_ints = (size_t)n_streams * 3 * height_mbs + DESC_RING * 6 * MAX_GROUPS;
    {
        const char *g = getenv("H264B2_GROUPS"), *gm = getenv("H264B2_GROUP_MIN");
        c->groups = g ? atoi(g) : 1;      // measured (S=128): 1 group 15.0k, 2 groups 14.0k, 4 groups 13.3k frames/s — the GPU is already issue-bound
        if (c->groups < 1) c->groups = 1;
        if (c->groups > MAX_GROUPS) c->groups = MAX_GROUPS;
        c->group_min = gm ? atoi(gm) : 8;
        if (c->group_min < 1) c->group_min = 1;
    }
    CK(cudaMalloc(&c->progress, c->progress_ints * 4));
    CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    { const char *e = getenv("H264B2_H2D_ZEROCOPY"); int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&c->st_h2d, cudaStreamNonBlocking, (e && atoi(e) > 0) ? hi : lo)); }
    CK(cudaStreamCreateWithFlags(&c->st_h2d2, cudaStreamNonBlocking));
    { const char *e = getenv("H264B2_H2D_STREAMS"); c->h2d_streams = (e && atoi(e) == 2) ? 2 : 1; }      // measured on the B200 box: 2 queues 10.8k frames/s e2e, 1 queue 11.1k
    { const char *e = getenv("H264B2_D2H_CHUNK_MB"); c->d2h_chunk = e && atoi(e) > 0 ? (size_t)atoi(e) << 20 : 0; }      // 0 = adaptive, see h264b2_read_pictures_async
    { const char *e = getenv("H264B2_D2H_ZEROCOPY"); c->d2h_ctas = e ? atoi(e) : 0; c->d2h_zerocopy = c->d2h_ctas > 0; }
    { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CK(cudaStreamCreateWithPriority(&c->st_d2h, cudaStreamNonBlocking, c->d2h_zerocopy ? hi : lo)); }
    { const char *e = getenv("H264B2_INTRA_SPLIT"); c->intra_split = !(e && atoi(e) == 0); }
    for (int i = 0; i < MAX_GROUPS; i++) {
        CK(cudaStreamCreateWithFlags(&c->st_chroma[i], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&c->chroma_fork[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->chroma_join[i], cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&c->st_g[i], cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c->st_side[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->side_fork[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->side_join[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
    CK(cudaHostAlloc(&c->h_desc, sizeof(PicDev) * n_streams * DESC_RING, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&c->h_desc_dev, c->h_desc, 0));
    { const char *e = getenv("H264B2_H2D_ZEROCOPY"); c->h2d_ctas = e ? atoi(e) : 0; }
    CK(cudaHostAlloc(&c->h_pull, sizeof(PullDesc) * n_streams * 8 * NSLOT, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&c->h_pull_dev, c->h_pull, 0));
    for (int i = 0; i < NOUT; i++) {
        CK(cudaHostAlloc(&c->h_snap[i], sizeof(uint8_t *) * n_streams * 2, cudaHostAllocMapped));      // [0, n): surfaces, [n, 2n): host destinations
        CK(cudaHostGetDevicePointer((void **)&c->h_snap_dev[i], c->h_snap[i], 0));
    }
    CK(cudaMalloc(&c->d_desc, sizeof(PicDev) * n_streams * DESC_RING));
    for (int i = 0; i < DESC_RING; i++) CK(cudaEventCreateWithFlags(&c->desc_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < NSLOT; i++) { CK(cudaEventCreateWithFlags(&c->h2d_done[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->h2d_done2[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->compute_done[i], cudaEventDisableTiming)); }
    for (int i = 0; i < NOUT; i++) { CK(cudaEventCreateWithFlags(&c->out_ready[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->out_done[i], cudaEventDisableTiming)); }
    CK(cudaMalloc(&c->d_ptrs, sizeof(uint8_t *) * n_streams));
    CK(cudaMalloc(&c->d_sums, 8 * n_streams));
    CK(cudaMallocHost(&c->h_sums, 8 * n_streams));
    CK(cudaMallocHost(&c->h_ptrs, sizeof(uint8_t *) * n_streams));
    CK(cudaEventCreate(&c->t0)); CK(cudaEventCreate(&c->t1));
    c->ev = (cudaEvent_t *)calloc(EV_POOL, sizeof(cudaEvent_t));
    if (init_tables(c)) return -10;
    // tensor maps of the decoded picture buffer for k_inter_tma (global strides must be multiples of 16 bytes: chroma rows need an even width in MBs)
    { const char *e = getenv("H264B2_INTER_V1"); c->inter_tma = !(e && atoi(e) != 0) && (width_mbs % 2 == 0) && width_mbs <= 256; }
    c->worklist_stride = ((size_t)c->nmb + 4 + 3) & ~(size_t)3;
    CK(cudaMalloc(&c->worklist, (size_t)n_streams * c->worklist_stride * sizeof(int)));
    CK(cudaMemset(c->worklist, 0, (size_t)n_streams * c->worklist_stride * sizeof(int)));
    if (c->inter_tma) {
        void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn || qr != cudaDriverEntryPointSuccess) { cudaGetLastError(); c->inter_tma = 0; }
        else {
            PFN_cuTensorMapEncodeTiled enc = (PFN_cuTensorMapEncodeTiled)fn;
            const cuuint64_t W = (cuuint64_t)width_mbs * 16, H = (cuuint64_t)height_mbs * 16, nsurf = (cuuint64_t)n_streams * surfaces_per_stream;
            const cuuint64_t dy[3] = { W, H, nsurf }, sy[2] = { W, c->frame_bytes };
            const cuuint32_t by[3] = { IT_LUMA_PITCH, IT_LUMA_ROWS, 1 }, e3[4] = { 1, 1, 1, 1 };
            const cuuint64_t dc[4] = { W / 2, H / 2, 2, nsurf }, sc[3] = { W / 2, (W / 2) * (H / 2), c->frame_bytes };
            const cuuint32_t bc[4] = { IT_CHROMA_PITCH, IT_CHROMA_ROWS, 2, 1 };
            CUresult r1 = enc(&c->map_y, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->surfaces, dy, sy, by, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult r2 = enc(&c->map_c, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, c->surfaces + W * H, dc, sc, bc, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return fail(-10, "cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
        }
    }
    c->trace_path = getenv("H264B2_TRACE");
    c->trace = c->trace_path != nullptr;
    return 0;
}

static void ev_destroy(cudaEvent_t e) { if (e) cudaEventDestroy(e); }
static void st_destroy(cudaStream_t st) { if (st) cudaStreamDestroy(st); }
extern "C" int h264b2_destroy(H264B2Context *c) {
    if (!c) return fail(-1, "null context");
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->bgr) cudaFree(c->bgr);
    cudaFree(c->worklist); cudaFree(c->surfaces); cudaFree(c->bs); cudaFree(c->res); cudaFree(c->progress); cudaFree(c->ls_flat); cudaFree(c->d_desc); cudaFreeHost(c->h_desc); cudaFreeHost(c->h_pull); for (int i = 0; i < NOUT; i++) cudaFreeHost(c->h_snap[i]);
    for (int i = 0; i < NSLOT; i++) { if (c->arena[i]) cudaFree(c->arena[i]); ev_destroy(c->h2d_done[i]); ev_destroy(c->h2d_done2[i]); ev_destroy(c->compute_done[i]); }
    for (int i = 0; i < NOUT; i++) { if (c->out_stage[i]) cudaFree(c->out_stage[i]); ev_destroy(c->out_ready[i]); ev_destroy(c->out_done[i]); }
    for (int i = 0; i < DESC_RING; i++) { ev_destroy(c->desc_ev[i]); ev_destroy(c->pre_done[i]); }
    ev_destroy(c->main_done[0]); ev_destroy(c->main_done[1]); st_destroy(c->st_pre);
    cudaFree(c->d_ptrs); cudaFree(c->d_sums); cudaFreeHost(c->h_sums); cudaFreeHost(c->h_ptrs);
    if (c->ev) { for (int i = 0; i < EV_POOL; i++) if (c->ev[i]) ev_destroy(c->ev[i]); }
    free(c->ev);
    ev_destroy(c->t0); ev_destroy(c->t1);
    st_destroy(c->st); st_destroy(c->st_h2d); st_destroy(c->st_h2d2); st_destroy(c->st_d2h);
    for (int i = 0; i < MAX_GROUPS; i++) { st_destroy(c->st_chroma[i]); ev_destroy(c->chroma_fork[i]); ev_destroy(c->chroma_join[i]); st_destroy(c->st_g[i]); st_destroy(c->st_side[i]); ev_destroy(c->join_ev[i]); ev_destroy(c->side_fork[i]); ev_destroy(c->side_join[i]); }
    ev_destroy(c->fork_ev);
    delete c;
    return 0;
}

// ------------------------------------------------------------------ timing helpers
static void class_begin(H264B2Context *c, int cls, cudaStream_t s) {
    if (!c->timing || c->ev_used + 2 > EV_POOL) return;
    for (int i = 0; i < 2; i++) if (!c->ev[c->ev_used + i]) cudaEventCreate(&c->ev[c->ev_used + i]);
    c->ev_class[c->ev_used / 2] = cls;
    cudaEventRecord(c->ev[c->ev_used], s);
}
static void class_end(H264B2Context *c, int cls, cudaStream_t s) {
    c->class_launches[cls]++;
    if (!c->timing || c->ev_used + 2 > EV_POOL) return;
    cudaEventRecord(c->ev[c->ev_used + 1], s);
    c->ev_used += 2;
}

static int validate(const H264B2Context *c, int n_pics, const int32_t *sids, const H264B2PicParams *pics) {
    if (!c || n_pics < 1 || n_pics > c->n_streams || !sids || !pics) return fail(-1, "submit: bad argument");
    std::vector<char> seen(c->n_streams, 0);
    for (int i = 0; i < n_pics; i++) {
        const H264B2PicParams &p = pics[i];
        if (sids[i] < 0 || sids[i] >= c->n_streams) return fail(-2, "submit: stream id %d out of range", sids[i]);
        if (seen[sids[i]]) return fail(-2, "submit: stream %d listed twice (pictures of one stream are serial)", sids[i]);
        seen[sids[i]] = 1;
        if (p.width_mbs != c->wmb || p.height_mbs != c->hmb) return fail(-3, "submit: picture is %dx%d MBs, context is %dx%d", p.width_mbs, p.height_mbs, c->wmb, c->hmb);
        if (p.dst_surface < 0 || p.dst_surface >= c->spp) return fail(-3, "submit: dst_surface %d out of range", p.dst_surface);
        if (p.mbaff_frame_flag && (c->hmb & 1)) return fail(-3, "submit: MBAFF picture with odd height in MBs");
        if (!p.mb_info || !p.intra_modes || !p.coef_offset || !p.weights || p.n_weights < 1) return fail(-3, "submit: missing array");
        if (p.has_inter && !p.motion) return fail(-3, "submit: has_inter without motion");
        if (p.n_coefs && !p.coefs) return fail(-3, "submit: n_coefs without coefs");
        if (p.packed & ~(H264B2_PACKED_COEFS | H264B2_PACKED_MOTION)) return fail(-3, "submit: unknown bits in packed (%d)", p.packed);
        if (p.custom_scaling && (!p.level_scale4 || !p.level_scale8)) return fail(-3, "submit: custom_scaling without tables");
    }
    return 0;
}

#define LCHK(name) do { if (c->debug_launch) { cudaError_t e_ = cudaPeekAtLastError(); if (e_ != cudaSuccess) return fail(-10, "launch of %s failed: %s (ng=%d np=%d nq=%d)", name, cudaGetErrorString(e_), ng, np, nq); } } while (0)
// enqueue the kernels of one batch; every array pointer in pics[] is a device pointer
static int launch_batch(H264B2Context *c, int n, const int32_t *sids, const H264B2PicParams *pics, const uint32_t *const *packed = nullptr, const uint32_t *const *packed_motion = nullptr,
                        cudaEvent_t in0 = nullptr, cudaEvent_t in1 = nullptr) {
    const int ring = c->desc_next; c->desc_next = (c->desc_next + 1) % DESC_RING;
    CK(cudaEventSynchronize(c->desc_ev[ring]));
    PicDev *hd = c->h_desc + (size_t)ring * c->n_streams, *dd = c->d_desc + (size_t)ring * c->n_streams;
    const int par = c->lookahead ? (int)(c->batch_no++ & 1) : 0;
    int any_inter = 0, any_deblock = 0, any_packed = 0, any_packed_motion = 0;
    // descriptors are laid out progressive pictures first, then "generic" ones (MBAFF, or wider than the fast paths
    // support): the wavefront kernels are launched once per kind, each with only its own code path compiled in
    std::vector<int> order; order.reserve(n);
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < n; i++) { const int gen = pics[i].mbaff_frame_flag || c->wmb > 256; if (gen == pass) order.push_back(i); }
    int n_prog = 0;
    for (int j = 0; j < n; j++) {
        const int i = order[j];
        const H264B2PicParams &p = pics[i];
        PicDev &d = hd[j];
        d.info = p.mb_info; d.modes = p.intra_modes; d.coef_off = p.coef_offset; d.motion = p.has_inter ? p.motion : nullptr;
        d.weights = p.weights; d.coefs = p.coefs; d.packed = packed ? packed[i] : nullptr; d.packed_motion = packed_motion ? packed_motion[i] : nullptr;
        any_packed |= d.packed != nullptr || d.packed_motion != nullptr; any_packed_motion |= d.packed_motion != nullptr;
        d.ls4 = p.custom_scaling ? p.level_scale4 : c->ls_flat;
        d.ls8 = p.custom_scaling ? p.level_scale8 : c->ls_flat + 2 * 2 * 6 * 16;
        d.stream_base = c->surfaces + (size_t)sids[i] * c->spp * c->frame_bytes;
        d.dst = (uint8_t *)d.stream_base + (size_t)p.dst_surface * c->frame_bytes;
        d.bs = c->bs + ((size_t)par * c->n_streams + sids[i]) * c->bs_stride;
        d.res = c->res + ((size_t)par * c->n_streams + sids[i]) * c->nmb * RES_MB_STRIDE;
        d.progress = c->progress + (size_t)sids[i] * 3 * c->hmb;
        d.frame_bytes = c->frame_bytes; d.surf0 = sids[i] * c->spp; d.worklist = c->worklist + (size_t)sids[i] * c->worklist_stride;
        d.wmb = c->wmb; d.hmb = c->hmb; d.mbaff = p.mbaff_frame_flag; d.cqp0 = p.chroma_qp_offset[0]; d.cqp1 = p.chroma_qp_offset[1];
        d.deblock_enable = p.deblock_enable; d.deblock_stop = p.deblock_stop_mb < c->nmb ? p.deblock_stop_mb : c->nmb;
        d.n_weights = p.n_weights; d.spp = c->spp; d.n_coefs = p.n_coefs;
        d.generic = p.mbaff_frame_flag || c->wmb > 256;
        n_prog += !d.generic;
        any_inter |= p.has_inter; any_deblock |= p.deblock_enable;
    }
    int *tickets = c->progress + (c->progress_ints - DESC_RING * 6 * MAX_GROUPS) + ring * 6 * MAX_GROUPS;
    const bool la = c->lookahead != 0;
    cudaStream_t sp = la ? c->st_pre : c->st;
    if (la) {
        if (in0) CK(cudaStreamWaitEvent(sp, in0, 0));
        if (in1) CK(cudaStreamWaitEvent(sp, in1, 0));
        CK(cudaStreamWaitEvent(sp, c->main_done[par], 0));       // the res / bs buffers of this parity are free (two batches ago)
    }
    class_begin(c, 0, sp);
    if (la) {
        k_prologue<<<1, 256, 0, sp>>>((const uint4 *)(c->h_desc_dev + (size_t)ring * c->n_streams), (uint4 *)dd, (int)(sizeof(PicDev) * n / 16), nullptr, 0);
        k_zero_ints<<<1, 256, 0, c->st>>>(c->progress, (int)c->progress_ints);
    } else
        k_prologue<<<1, 256, 0, c->st>>>((const uint4 *)(c->h_desc_dev + (size_t)ring * c->n_streams), (uint4 *)dd, (int)(sizeof(PicDev) * n / 16), c->progress, (int)c->progress_ints);
    CK(cudaEventRecord(c->desc_ev[ring], sp));
    for (int j = 0; j < n; j++) if (pics[order[j]].clear_surface) CK(cudaMemsetAsync(hd[j].dst, 0, c->frame_bytes, c->st));
    if (any_packed) k_expand<<<dim3(32, n, any_packed_motion ? 2 : 1), 256, 0, sp>>>(dd);
    if (any_packed_motion) k_unmotion<<<dim3((c->nmb * 2 + 255) / 256, n), 256, 0, sp>>>(dd);
    class_end(c, 0, sp);
    if (la) {
        class_begin(c, 5, sp);
        k_residual<<<dim3((c->nmb + 7) / 8, n), 256, 0, sp>>>(dd);
        class_end(c, 5, sp);
        if (any_deblock) {
            class_begin(c, 3, sp);
            if (n_prog) { if (c->deblock_v1) k_bs_prog<<<dim3((c->wmb + 7) / 8, c->hmb, n_prog), 256, 0, sp>>>(dd); else k_bs_prog2<<<dim3((c->nmb + 127) / 128, n_prog), 128, 0, sp>>>(dd); }
            if (n - n_prog) k_bs<<<dim3((c->nmb + 7) / 8, n - n_prog), 256, 0, sp>>>(dd + n_prog);
            class_end(c, 3, sp);
        }
        CK(cudaEventRecord(c->pre_done[ring], sp));
        CK(cudaStreamWaitEvent(c->st, c->pre_done[ring], 0));
    }
    int G = c->groups;
    while (G > 1 && n / G < c->group_min) G--;
    const int bands = (c->hmb + WF_ROWS - 1) / WF_ROWS;
    if (G > 1) CK(cudaEventRecord(c->fork_ev, c->st));
    for (int g = 0; g < G; g++) {
        const int b0 = (int)((long long)n * g / G), b1 = (int)((long long)n * (g + 1) / G), ng = b1 - b0;
        if (ng <= 0) continue;
        cudaStream_t sg = G > 1 ? c->st_g[g] : c->st;
        if (G > 1) CK(cudaStreamWaitEvent(sg, c->fork_ev, 0));
        int g_inter = 0, g_deblock = 0;
        for (int j = b0; j < b1; j++) { g_inter |= pics[order[j]].has_inter; g_deblock |= pics[order[j]].deblock_enable; }
        const int np = std::max(0, std::min(b1, n_prog) - b0), nq = ng - np;      // progressive / generic pictures of this group
        const PicDev *dg = dd + b0;
        if (!la) {
            class_begin(c, 5, sg);
            k_residual<<<dim3((c->nmb + 7) / 8, ng), 256, 0, sg>>>(dg);
            class_end(c, 5, sg);
        }
        if (g_inter) {
            class_begin(c, 1, sg);
            // progressive pictures (first np descriptors): macroblocks with one vector per list and windows inside the picture go through
            // the TMA-staged kernel, k_inter does the rest (and everything of MBAFF pictures); the two write disjoint macroblocks
            if (c->inter_tma && np > 0) {
                k_zero_worklists<<<1, 256, 0, sg>>>(c->worklist, c->n_streams, c->worklist_stride);
                k_inter_tma<<<dim3((c->wmb + 4 * IT_MBS - 1) / (4 * IT_MBS), c->hmb, np), 128, 0, sg>>>(dg, c->map_y, c->map_c);
                LCHK("k_inter_tma");
                k_inter_list<<<dim3(IT_LIST_CTAS, np), 128, 0, sg>>>(dg);
                LCHK("k_inter_list");
                if (nq) k_inter<<<dim3((c->wmb + 3) / 4, c->hmb, nq), 128, 0, sg>>>(dg + np, 0);
            } else k_inter<<<dim3((c->wmb + 3) / 4, c->hmb, ng), 128, 0, sg>>>(dg, 0);
            LCHK("k_inter");
            class_end(c, 1, sg);
        }
        // boundary strengths need only side info, not samples: compute them before the wavefronts so that the
        // intra -> deblock chains of the progressive and of the generic (MBAFF) pictures can run side by side
        if (g_deblock && !la) {
            class_begin(c, 3, sg);
            if (np) { if (c->deblock_v1) k_bs_prog<<<dim3((c->wmb + 7) / 8, c->hmb, np), 256, 0, sg>>>(dg); else k_bs_prog2<<<dim3((c->nmb + 127) / 128, np), 128, 0, sg>>>(dg); }      // progressive pictures come first in dg
            if (nq) k_bs<<<dim3((c->nmb + 7) / 8, nq), 256, 0, sg>>>(dg + np);
            class_end(c, 3, sg);
        }
        cudaStream_t sq = sg;
        if (np && nq) {                 // generic pictures continue on a side stream
            sq = c->st_side[g];
            CK(cudaEventRecord(c->side_fork[g], sg));
            CK(cudaStreamWaitEvent(sq, c->side_fork[g], 0));
        }
        if (np) {
            class_begin(c, 2, sg);
            if (c->intra_split) {
                // luma and chroma intra prediction as two independent wavefronts: side by side on two streams when the look-ahead schedule is
                // on, one after the other when every kernel is timed alone
                cudaStream_t sc = la ? c->st_chroma[g] : sg;
                if (la) { CK(cudaEventRecord(c->chroma_fork[g], sg)); CK(cudaStreamWaitEvent(sc, c->chroma_fork[g], 0)); }
                k_intra<false, 2><<<np * bands, WF_THREADS, 0, sc>>>(dg, np, bands, tickets + 6 * g + 4);
                if (la) CK(cudaEventRecord(c->chroma_join[g], sc));
                k_intra<false, 1><<<np * bands, WF_THREADS, 0, sg>>>(dg, np, bands, tickets + 6 * g);
                if (la) CK(cudaStreamWaitEvent(sg, c->chroma_join[g], 0));
            } else k_intra<false, 3><<<np * bands, WF_THREADS, 0, sg>>>(dg, np, bands, tickets + 6 * g);
            LCHK("k_intra<false>");
            class_end(c, 2, sg);
            if (g_deblock) {
                class_begin(c, 4, sg);
                if (c->deblock_v1) k_deblock<false><<<np * bands, WF_THREADS, 0, sg>>>(dg, np, bands, tickets + 6 * g + 2);
                else { const int bands3 = (c->hmb + DB3_ROWS - 1) / DB3_ROWS; k_deblock3<<<((np + DB_PPW - 1) / DB_PPW) * bands3, DB3_THREADS, 0, sg>>>(dg, np, bands3, tickets + 6 * g + 2); }
                class_end(c, 4, sg);
            }
        }
        if (nq) {
            class_begin(c, 2, sq);
            k_intra<true><<<nq * bands, WF_THREADS, 0, sq>>>(dg + np, nq, bands, tickets + 6 * g + 1);
            LCHK("k_intra<true>");
            class_end(c, 2, sq);
            if (g_deblock) {
                class_begin(c, 4, sq);
                k_deblock<true><<<nq * bands, WF_THREADS, 0, sq>>>(dg + np, nq, bands, tickets + 6 * g + 3);
                LCHK("k_deblock<true>");
                class_end(c, 4, sq);
            }
            if (sq != sg) { CK(cudaEventRecord(c->side_join[g], sq)); CK(cudaStreamWaitEvent(sg, c->side_join[g], 0)); }
        }
        if (G > 1) { CK(cudaEventRecord(c->join_ev[g], sg)); CK(cudaStreamWaitEvent(c->st, c->join_ev[g], 0)); }
    }
    (void)any_inter; (void)any_deblock;
    if (la) CK(cudaEventRecord(c->main_done[par], c->st));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int h264b2_submit_device(H264B2Context *c, int n_pics, const int32_t *sids, const H264B2PicParams *pics) {
    int r = validate(c, n_pics, sids, pics);
    if (r) return r;
    for (int i = 0; i < n_pics; i++) if (pics[i].packed) return fail(-3, "submit_device: packed arrays travel through h264b2_submit (host arrays)");
    CK(cudaSetDevice(c->device));
    return launch_batch(c, n_pics, sids, pics);
}

// ---- packed coefficient transport, host side (layout: see k_expand)
extern "C" size_t h264b2_pack_coefs_bound(uint32_t n_coefs) {
    const size_t nc = ((size_t)n_coefs + 15) / 16, ng = (nc + 31) / 32;
    return (16 + ng * 4 + ((nc + 1) & ~(size_t)1) * 2 + nc * 32 + 15) & ~(size_t)15;
}
// streaming writer of the blob: chunks of 16 levels arrive in order (the motion packer produces them four records at a time)
struct BlobWriter {
    uint32_t *hdr, *base; uint16_t *map; int16_t *val; size_t nc, fixed, vcap, nv = 0, ch = 0; bool overflow = false;
    bool begin(void *out, size_t cap, size_t n_levels) {
        nc = (n_levels + 15) / 16;
        const size_t ng = (nc + 31) / 32;
        fixed = 16 + ng * 4 + ((nc + 1) & ~(size_t)1) * 2;
        if (cap < fixed) return false;
        hdr = (uint32_t *)out; base = hdr + 4; map = (uint16_t *)(base + ng); val = (int16_t *)((uint8_t *)out + fixed);
        vcap = (cap - fixed) / 2;
        if (nc & 1) map[nc] = 0;
        return true;
    }
    inline void chunk(const int16_t *c, size_t n) {          // n = 16 except for the tail chunk
        if ((ch & 31) == 0) base[ch >> 5] = (uint32_t)nv;
        if (n == 16) {                                       // most chunks hold nothing: test 32 bytes at once
            uint64_t q[4]; memcpy(q, c, 32);
            if (!(q[0] | q[1] | q[2] | q[3])) { map[ch++] = 0; return; }
        }
        if (nv + 16 > vcap) {                                // exact check only when space is short
            size_t cnt = 0; for (size_t k = 0; k < n; k++) cnt += c[k] != 0;
            if (nv + cnt > vcap) { overflow = true; map[ch++] = 0; return; }
        }
        uint32_t bm = 0;
        for (size_t k = 0; k < n; k++) { const int16_t v = c[k]; val[nv] = v; const unsigned nz = v != 0; bm |= nz << k; nv += nz; }
        map[ch++] = (uint16_t)bm;
    }
    int end(size_t cap, size_t *bytes) {
        if (overflow) return -3;
        size_t total = fixed + nv * 2;
        while (total & 15) { if (total + 2 <= cap) *(int16_t *)((uint8_t *)hdr + total) = 0; total += 2; }
        if (total > cap) return -3;
        hdr[0] = H264B2_PACK_MAGIC; hdr[1] = (uint32_t)nc; hdr[2] = (uint32_t)nv; hdr[3] = (uint32_t)total;
        *bytes = total;
        return 0;
    }
};
extern "C" int h264b2_pack_coefs(const int16_t *dense, uint32_t n_coefs, void *out, size_t cap, size_t *bytes) {
    if ((!dense && n_coefs) || !out || !bytes || ((uintptr_t)out & 3)) return fail(-1, "pack_coefs: bad argument");
    BlobWriter w;
    if (!w.begin(out, cap, n_coefs)) return fail(-3, "pack_coefs: output buffer too small");
    for (size_t lo = 0; lo < n_coefs; lo += 16) w.chunk(dense + lo, std::min<size_t>(16, n_coefs - lo));
    return w.end(cap, bytes) ? fail(-3, "pack_coefs: output buffer too small") : 0;
}
extern "C" int h264b2_unpack_coefs(const void *packed, int16_t *dense, uint32_t n_coefs) {
    const uint32_t *hdr = (const uint32_t *)packed;
    if (!packed || (!dense && n_coefs) || hdr[0] != H264B2_PACK_MAGIC || hdr[1] != (n_coefs + 15) / 16) return fail(-3, "unpack_coefs: malformed blob");
    const size_t nc = hdr[1], ng = (nc + 31) / 32;
    const uint16_t *map = (const uint16_t *)(hdr + 4 + ng);
    const int16_t *val = (const int16_t *)(map + ((nc + 1) & ~(size_t)1));
    size_t nv = 0;
    for (size_t k = 0; k < n_coefs; k++) dense[k] = ((map[k >> 4] >> (k & 15)) & 1) ? val[nv++] : (int16_t)0;
    return nv == hdr[2] ? 0 : fail(-3, "unpack_coefs: value count mismatch");
}

extern "C" int h264b2_pack_motion(const H264B2MbMotion *motion, uint32_t n_mbs, void *out, size_t cap, size_t *bytes) {
    if ((!motion && n_mbs) || !out || !bytes || ((uintptr_t)out & 3)) return fail(-1, "pack_motion: bad argument");
    constexpr size_t W = sizeof(H264B2MbMotion) / 4, H16 = sizeof(H264B2MbMotion) / 2;      // 38 words = 76 int16 per record
    BlobWriter w;
    if (!w.begin(out, cap, (size_t)n_mbs * H16)) return fail(-3, "pack_motion: output buffer too small");
    // four records = 304 int16 = 19 whole chunks: XOR-chain them in a small buffer and hand the chunks on
    uint32_t buf[4 * W];
    for (size_t a = 0; a < n_mbs; a += 4) {
        const size_t n = std::min<size_t>(4, n_mbs - a);
        memcpy(buf, motion + a, n * sizeof(H264B2MbMotion));
        for (size_t k = 0; k < n; k++)
            for (int l = 0; l < 2; l++) { uint32_t *q = buf + k * W + l * 16; for (int r = 15; r >= 1; r--) q[r] ^= q[r - 1]; }
        const int16_t *c = (const int16_t *)buf;
        for (size_t lo = 0; lo < n * H16; lo += 16) w.chunk(c + lo, std::min<size_t>(16, n * H16 - lo));
    }
    return w.end(cap, bytes) ? fail(-3, "pack_motion: output buffer too small") : 0;
}
extern "C" int h264b2_unpack_motion(const void *packed, H264B2MbMotion *motion, uint32_t n_mbs) {
    int r = h264b2_unpack_coefs(packed, (int16_t *)motion, (uint32_t)(n_mbs * (sizeof(H264B2MbMotion) / 2)));
    if (r) return r;
    for (size_t a = 0; a < n_mbs; a++)
        for (int l = 0; l < 2; l++) { uint32_t *w = (uint32_t *)&motion[a].mv[l][0][0]; for (int i = 1; i < 16; i++) w[i] ^= w[i - 1]; }
    return 0;
}

// ---- host-array submit: one DMA per contiguous span, straight from the caller's memory (pinned memory
// from h264b2_host_alloc makes the copies asynchronous; pageable memory works but is staged by the driver).
struct Span { const uint8_t *p; size_t n; const void **slot; };
// header of a packed blob: magic, chunk count as expected, and a total size that matches the value count
static inline bool blob_header_ok(const uint32_t *hdr, uint32_t want_chunks) {
    if (((uintptr_t)hdr & 15) || hdr[0] != H264B2_PACK_MAGIC || hdr[1] != want_chunks) return false;
    const size_t nc = hdr[1], ng = (nc + 31) / 32, fixed = 16 + ng * 4 + ((nc + 1) & ~(size_t)1) * 2;
    return hdr[2] <= nc * 16 && hdr[3] == ((fixed + (size_t)hdr[2] * 2 + 15) & ~(size_t)15);
}
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int h264b2_submit(H264B2Context *c, int n_pics, const int32_t *sids, const H264B2PicParams *pics) {
    int r = validate(c, n_pics, sids, pics);
    if (r) return r;
    CK(cudaSetDevice(c->device));
    // every check that can fail comes BEFORE an arena slot is taken: a rejected submit must not shift the H264B2_SUBMIT_DEPTH contract
    for (int i = 0; i < n_pics; i++) {
        const H264B2PicParams &p = pics[i];
        const size_t nmb = (size_t)c->nmb;
        if ((p.packed & H264B2_PACKED_MOTION) && p.has_inter && !blob_header_ok((const uint32_t *)p.motion, (uint32_t)((nmb * (sizeof(H264B2MbMotion) / 2) + 15) / 16)))
            return fail(-3, "submit: malformed packed motion blob (picture %d)", i);
        if ((p.packed & H264B2_PACKED_COEFS) && p.n_coefs && !blob_header_ok((const uint32_t *)p.coefs, (p.n_coefs + 15) / 16))
            return fail(-3, "submit: malformed packed coefficient blob (picture %d)", i);
    }
    const int slot = c->slot_next; c->slot_next = (c->slot_next + 1) % NSLOT;
    CK(cudaEventSynchronize(c->h2d_done[slot]));      // the host arrays of the submit NSLOT calls ago have left (API contract)
    if (c->h2d_streams > 1) CK(cudaEventSynchronize(c->h2d_done2[slot]));
    std::vector<H264B2PicParams> dev(pics, pics + n_pics);
    // plan
    size_t need = 0;
    std::vector<Span> spans; spans.reserve((size_t)n_pics * 8);
    std::vector<const uint32_t *> packed((size_t)n_pics, nullptr), packed_m((size_t)n_pics, nullptr);
    bool any_packed = false, any_packed_m = false;
    for (int i = 0; i < n_pics; i++) {
        H264B2PicParams &p = dev[i];
        const size_t nmb = (size_t)c->nmb;
        size_t coef_bytes = (size_t)p.n_coefs * 2;
        size_t motion_bytes = p.has_inter ? nmb * sizeof(H264B2MbMotion) : 0;
        if ((p.packed & H264B2_PACKED_MOTION) && p.has_inter) {
            const uint32_t *hdr = (const uint32_t *)p.motion;
            const uint32_t want = (uint32_t)((nmb * (sizeof(H264B2MbMotion) / 2) + 15) / 16);
            if (!blob_header_ok(hdr, want)) return fail(-3, "submit: malformed packed motion blob (picture %d)", i);
            motion_bytes = hdr[3];
            need += al256((size_t)want * 32) + 256;
            any_packed_m = true;
        }
        if ((p.packed & H264B2_PACKED_COEFS) && p.n_coefs) {
            const uint32_t *hdr = (const uint32_t *)p.coefs;
            if (!blob_header_ok(hdr, (p.n_coefs + 15) / 16)) return fail(-3, "submit: malformed packed coefficient blob (picture %d)", i);
            coef_bytes = hdr[3];
            need += al256((size_t)hdr[1] * 32) + 256;      // the dense array k_expand rebuilds
            any_packed = true;
        }
        Span s[8] = {
            { (const uint8_t *)p.mb_info, nmb * sizeof(H264B2MbInfo), (const void **)&p.mb_info },
            { (const uint8_t *)p.intra_modes, nmb * 8, (const void **)&p.intra_modes },
            { (const uint8_t *)p.coef_offset, nmb * 4, (const void **)&p.coef_offset },
            { (const uint8_t *)(p.has_inter ? p.motion : nullptr), motion_bytes, (const void **)&p.motion },
            { (const uint8_t *)p.weights, (size_t)p.n_weights * sizeof(H264B2Weight), (const void **)&p.weights },
            { (const uint8_t *)p.coefs, coef_bytes, (const void **)&p.coefs },
            { (const uint8_t *)(p.custom_scaling ? p.level_scale4 : nullptr), p.custom_scaling ? (size_t)2 * 2 * 6 * 16 * 2 : 0, (const void **)&p.level_scale4 },
            { (const uint8_t *)(p.custom_scaling ? p.level_scale8 : nullptr), p.custom_scaling ? (size_t)2 * 2 * 6 * 64 * 2 : 0, (const void **)&p.level_scale8 },
        };
        for (int k = 0; k < 8; k++) { if (s[k].n) { spans.push_back(s[k]); need += al256(s[k].n) + 256; } else *s[k].slot = nullptr; }
    }
    if (c->arena_cap[slot] < need) {
        CK(cudaEventSynchronize(c->compute_done[slot]));
        if (c->arena[slot]) CK(cudaFree(c->arena[slot]));
        c->arena_cap[slot] = need + need / 4;
        CK(cudaMalloc(&c->arena[slot], c->arena_cap[slot]));
    }
    CK(cudaStreamWaitEvent(c->st_h2d, c->compute_done[slot], 0));
    if (c->h2d_streams > 1) CK(cudaStreamWaitEvent(c->st_h2d2, c->compute_done[slot], 0));
    trace_mark(c, 0, c->st_h2d);
    // copy: merge spans that are adjacent in host memory (same relative alignment kept)
    size_t off = 0;
    size_t i = 0;
    unsigned ncopy = 0;
    PullDesc *pull = c->h_pull + (size_t)slot * c->n_streams * 8;
    int npull = 0;
    bool zc = c->h2d_ctas > 0;
    while (i < spans.size()) {
        size_t j = i;
        const uint8_t *lo = spans[i].p; const uint8_t *hi = lo + spans[i].n;
        // the kernels use 16-byte loads on these arrays (header: alignment contract).  A host array that does not start on a 16-byte
        // boundary (e.g. a numpy view into a file image) travels alone and lands on a 256-byte boundary of the arena instead.
        const bool lo_ok = ((uintptr_t)lo & 15) == 0;
        while (lo_ok && j + 1 < spans.size() && spans[j + 1].p >= hi && (size_t)(spans[j + 1].p - hi) <= 64 && ((uintptr_t)spans[j + 1].p & 15) == 0) { j++; hi = spans[j].p + spans[j].n; }
        // keep the host address's offset within 256 so that every array keeps its natural alignment
        off = al256(off) + (lo_ok ? ((size_t)(uintptr_t)lo & 255) : 0);
        // picture-sized DMAs leave gaps on one copy queue (profiles/r01_pcie_probe.txt): alternate between two
        cudaPointerAttributes at;
        if (zc && lo_ok && cudaPointerGetAttributes(&at, lo) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            pull[npull].src = (const uint8_t *)at.devicePointer; pull[npull].dst = c->arena[slot] + off; pull[npull].bytes = (size_t)(hi - lo); npull++;
        } else {
            if (zc) cudaGetLastError();
            CK(cudaMemcpyAsync(c->arena[slot] + off, lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, (c->h2d_streams > 1 && (ncopy & 1)) ? c->st_h2d2 : c->st_h2d));
        }
        ncopy++;
        for (size_t k = i; k <= j; k++) *spans[k].slot = c->arena[slot] + off + (spans[k].p - lo);
        off += (size_t)(hi - lo);
        i = j + 1;
    }
    if (any_packed || any_packed_m) for (int k = 0; k < n_pics; k++) {
        H264B2PicParams &p = dev[k];
        if ((p.packed & H264B2_PACKED_COEFS) && p.n_coefs) {
            packed[k] = (const uint32_t *)p.coefs;                 // the blob's device copy
            off = al256(off);
            p.coefs = (const int16_t *)(c->arena[slot] + off);      // where k_expand writes the dense array
            off += (size_t)((p.n_coefs + 15) / 16) * 32;
        }
        if ((p.packed & H264B2_PACKED_MOTION) && p.has_inter) {
            packed_m[k] = (const uint32_t *)p.motion;
            off = al256(off);
            p.motion = (const H264B2MbMotion *)(c->arena[slot] + off);
            off += (((size_t)c->nmb * (sizeof(H264B2MbMotion) / 2) + 15) / 16) * 32;
        }
    }
    if (npull) { k_pull<<<c->h2d_ctas, 256, 0, c->st_h2d>>>(c->h_pull_dev + (size_t)slot * c->n_streams * 8, npull); CK(cudaGetLastError()); }
    c->h2d_last_bytes = off; c->h2d_last_copies = ncopy;
    CK(cudaEventRecord(c->h2d_done[slot], c->st_h2d));
    trace_mark(c, 1, c->st_h2d);
    CK(cudaStreamWaitEvent(c->st, c->h2d_done[slot], 0));
    if (c->h2d_streams > 1) { CK(cudaEventRecord(c->h2d_done2[slot], c->st_h2d2)); CK(cudaStreamWaitEvent(c->st, c->h2d_done2[slot], 0)); }
    trace_mark(c, 2, c->st);
    r = launch_batch(c, n_pics, sids, dev.data(), any_packed ? packed.data() : nullptr, any_packed_m ? packed_m.data() : nullptr,
                     c->h2d_done[slot], c->h2d_streams > 1 ? c->h2d_done2[slot] : nullptr);
    if (r) return r;
    CK(cudaEventRecord(c->compute_done[slot], c->st));
    trace_mark(c, 3, c->st);
    if (c->trace && c->trace_n < 256) c->trace_n++;
    return 0;
}

// ------------------------------------------------------------------ surfaces
static int surf_ptr(H264B2Context *c, int sid, int surface, uint8_t **p) {
    if (!c || sid < 0 || sid >= c->n_streams || surface < 0 || surface >= c->spp) return fail(-2, "stream/surface out of range");
    *p = c->surfaces + ((size_t)sid * c->spp + surface) * c->frame_bytes;
    return 0;
}
extern "C" int h264b2_surface_ptr(H264B2Context *c, int sid, int surface, void **dev_ptr) {
    uint8_t *p; int r = surf_ptr(c, sid, surface, &p); if (r) return r; *dev_ptr = p; return 0;
}
extern "C" int h264b2_debug_intra_tables(int n, uint16_t *out) {
    if ((n != 4 && n != 8) || !out) return fail(-1, "debug_intra_tables: n must be 4 or 8");
    fill_intra_tables(n, out);
    return 0;
}
extern "C" int h264b2_set_lookahead(H264B2Context *c, int on) {
    if (!c) return fail(-1, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st_pre)); CK(cudaStreamSynchronize(c->st));
    c->lookahead = on != 0;
    return 0;
}
extern "C" int h264b2_sync(H264B2Context *c) {
    if (!c) return fail(-1, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st_h2d)); CK(cudaStreamSynchronize(c->st_h2d2)); CK(cudaStreamSynchronize(c->st_pre)); CK(cudaStreamSynchronize(c->st)); CK(cudaStreamSynchronize(c->st_d2h));
    trace_dump(c);
    return 0;
}
extern "C" int h264b2_read_picture(H264B2Context *c, int sid, int surface, uint8_t *host) {
    uint8_t *p; int r = surf_ptr(c, sid, surface, &p); if (r) return r;
    if (!host) return fail(-1, "null host pointer");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(host, p, c->frame_bytes, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}
extern "C" int h264b2_write_picture(H264B2Context *c, int sid, int surface, const uint8_t *host) {
    uint8_t *p; int r = surf_ptr(c, sid, surface, &p); if (r) return r;
    if (!host) return fail(-1, "null host pointer");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(p, host, c->frame_bytes, cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}
// Asynchronous read-back of n surfaces (n <= n_streams) into host memory (pinned for true overlap): the
// surfaces are snapshotted device-to-device on the launch stream, then drained to the host on a separate
// stream so the next submit does not wait for PCIe.  Complete after h264b2_sync().
extern "C" int h264b2_read_pictures_async(H264B2Context *c, int n, const int32_t *sids, const int32_t *surfaces, uint8_t *const *host) {
    if (!c || n < 1 || n > c->n_streams || !sids || !surfaces || !host) return fail(-1, "read_pictures_async: bad argument");
    CK(cudaSetDevice(c->device));
    const int s = c->out_next; c->out_next = (c->out_next + 1) % NOUT;
    const size_t need = (size_t)n * c->frame_bytes;
    if (c->out_cap[s] < need) {
        CK(cudaEventSynchronize(c->out_done[s]));
        if (c->out_stage[s]) CK(cudaFree(c->out_stage[s]));
        c->out_cap[s] = (size_t)c->n_streams * c->frame_bytes;
        CK(cudaMalloc(&c->out_stage[s], c->out_cap[s]));
    }
    CK(cudaEventSynchronize(c->out_done[s]));     // the pointer list of this slot is free again (2 read-backs ago)
    for (int i = 0; i < n; i++) { int r = surf_ptr(c, sids[i], surfaces[i], &c->h_snap[s][i]); if (r) return r; }
    k_snapshot<<<dim3(32, n), 256, 0, c->st>>>(c->h_snap_dev[s], c->out_stage[s], c->frame_bytes / 16);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->out_ready[s], c->st));
    CK(cudaStreamWaitEvent(c->st_d2h, c->out_ready[s], 0));
    trace_mark(c, 4, c->st_d2h, 1);
    if (c->d2h_zerocopy) {
        bool ok = true;
        for (int i = 0; i < n && ok; i++) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, host[i]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) { cudaGetLastError(); ok = false; break; }
            c->h_snap[s][c->n_streams + i] = (uint8_t *)at.devicePointer;
        }
        if (ok) {
            k_drain<<<c->d2h_ctas, 256, 0, c->st_d2h>>>(c->out_stage[s], c->h_snap_dev[s] + c->n_streams, c->frame_bytes / 16, n);
            CK(cudaGetLastError());
            CK(cudaEventRecord(c->out_done[s], c->st_d2h));
            return 0;
        }
    }
    // the staging buffer is contiguous: destinations that are adjacent in host memory leave as ONE transfer (bounded chunks)
    // Policy (profiles/r01_pcie_probe2.txt, B200 box): with picture-sized H2D transfers in flight a merged read-back starves
    // them (e2e 11.1k -> 9.8k frames/s); when the submits arrive as batch-sized transfers, merging the read-back as well lifts
    // both directions to ~50 GB/s.  So: merge iff the last submit averaged >= 16 MB per transfer (or H264B2_D2H_CHUNK_MB says so).
    size_t chunk = c->d2h_chunk;
    if (!chunk) chunk = (c->h2d_last_copies && c->h2d_last_bytes / c->h2d_last_copies >= ((size_t)16 << 20)) ? (size_t)1 << 30 : 1;
    const int per = (int)std::max<size_t>(1, chunk / c->frame_bytes);
    for (int i = 0; i < n; ) {
        int j = i + 1;
        while (j < n && j - i < per && host[j] == host[j - 1] + c->frame_bytes) j++;
        CK(cudaMemcpyAsync(host[i], c->out_stage[s] + (size_t)i * c->frame_bytes, (size_t)(j - i) * c->frame_bytes, cudaMemcpyDeviceToHost, c->st_d2h));
        i = j;
    }
    trace_mark(c, 5, c->st_d2h, 1);
    CK(cudaEventRecord(c->out_done[s], c->st_d2h));
    return 0;
}
extern "C" int h264b2_checksum_pictures(H264B2Context *c, int n, const int32_t *sids, const int32_t *surfaces, uint64_t *sums) {
    if (!c || n < 1 || n > c->n_streams || !sids || !surfaces || !sums) return fail(-1, "checksum_pictures: bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st));      // h_ptrs / h_sums are reused
    for (int i = 0; i < n; i++) { int r = surf_ptr(c, sids[i], surfaces[i], &c->h_ptrs[i]); if (r) return r; }
    CK(cudaMemcpyAsync(c->d_ptrs, c->h_ptrs, sizeof(uint8_t *) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemsetAsync(c->d_sums, 0, 8 * n, c->st));
    k_checksum<<<dim3(64, n), 256, 0, c->st>>>(c->d_ptrs, c->frame_bytes / 4, c->d_sums);
    CK(cudaMemcpyAsync(c->h_sums, c->d_sums, 8 * n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    for (int i = 0; i < n; i++) sums[i] = c->h_sums[i];
    return 0;
}
extern "C" int h264b2_checksum_picture(H264B2Context *c, int sid, int surface, uint64_t *sum) {
    const int32_t s = sid, f = surface;
    return h264b2_checksum_pictures(c, 1, &s, &f, sum);
}

// Output stage: convert a reconstructed surface to BGR24 on the GPU (what the reference's BMP writer and player
// do on the CPU, H264PictureBase.cpp:525-548) and copy it to the host.  width_bytes >= 3*W (BMP rows are padded to
// 4 bytes); flip_lines = 1 writes bottom-up like convertYuv420pToBgr24FlipLines.
extern "C" int h264b2_read_picture_bgr24(H264B2Context *c, int sid, int surface, uint8_t *host_bgr24, int width_bytes, int flip_lines) {
    uint8_t *p; int r = surf_ptr(c, sid, surface, &p); if (r) return r;
    const int W = c->wmb * 16, H = c->hmb * 16;
    if (!host_bgr24 || width_bytes < 3 * W) return fail(-1, "read_picture_bgr24: bad argument");
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)width_bytes * H;
    if (c->bgr_cap < bytes) { if (c->bgr) CK(cudaFree(c->bgr)); CK(cudaMalloc(&c->bgr, bytes)); c->bgr_cap = bytes; CK(cudaMemset(c->bgr, 0, bytes)); }
    k_bgr24<<<dim3((W / 4 + 127) / 128, H), 128, 0, c->st>>>(p, W, H, c->bgr, width_bytes, flip_lines);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host_bgr24, c->bgr, bytes, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

// ------------------------------------------------------------------ memory helpers
extern "C" int h264b2_dev_alloc(H264B2Context *c, size_t bytes, void **p) { if (!c || !p) return fail(-1, "bad argument"); CK(cudaSetDevice(c->device)); CK(cudaMalloc(p, bytes ? bytes : 1)); return 0; }
extern "C" int h264b2_dev_free(H264B2Context *c, void *p) { if (!c) return fail(-1, "bad argument"); CK(cudaSetDevice(c->device)); CK(cudaFree(p)); return 0; }
extern "C" int h264b2_dev_upload(H264B2Context *c, void *dst, const void *src, size_t bytes) {
    if (!c || !dst || !src) return fail(-1, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int h264b2_dev_copy(H264B2Context *c, void *dst, const void *src, size_t bytes) {
    if (!c || !dst || !src) return fail(-1, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
    return 0;
}
extern "C" int h264b2_host_alloc(H264B2Context *c, size_t bytes, void **p) { if (!c || !p) return fail(-1, "bad argument"); CK(cudaSetDevice(c->device)); CK(cudaMallocHost(p, bytes ? bytes : 1)); return 0; }
extern "C" int h264b2_host_free(H264B2Context *c, void *p) { if (!c) return fail(-1, "bad argument"); CK(cudaSetDevice(c->device)); CK(cudaFreeHost(p)); return 0; }

// ------------------------------------------------------------------ timing
extern "C" int h264b2_timer_start(H264B2Context *c) {
    if (!c) return fail(-1, "null context");
    CK(cudaSetDevice(c->device));
    c->timing = 1; c->ev_used = 0;
    for (int i = 0; i < KCLASSES; i++) { c->class_ms[i] = 0; c->class_launches[i] = 0; }
    CK(cudaEventRecord(c->t0, c->st));
    return 0;
}
extern "C" int h264b2_timer_stop(H264B2Context *c, float *ms) {
    if (!c || !ms) return fail(-1, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st_h2d)); CK(cudaStreamSynchronize(c->st_h2d2));
    for (int i = 0; i < NOUT; i++) CK(cudaStreamWaitEvent(c->st, c->out_done[i], 0));   // t1 also covers pending read-backs
    CK(cudaEventRecord(c->t1, c->st));
    CK(cudaEventSynchronize(c->t1));
    CK(cudaStreamSynchronize(c->st_d2h));
    CK(cudaEventElapsedTime(ms, c->t0, c->t1));
    for (int i = 0; i + 1 < c->ev_used; i += 2) {
        float t = 0; CK(cudaEventElapsedTime(&t, c->ev[i], c->ev[i + 1]));
        c->class_ms[c->ev_class[i / 2]] += t;
    }
    c->timing = 0;
    return 0;
}
extern "C" int h264b2_kernel_times(H264B2Context *c, float *ms5, int64_t *launches) {
    if (!c || !ms5) return fail(-1, "bad argument");
    for (int i = 0; i < KCLASSES; i++) { ms5[i] = c->class_ms[i]; if (launches) launches[i] = c->class_launches[i]; }
    return 0;
}
