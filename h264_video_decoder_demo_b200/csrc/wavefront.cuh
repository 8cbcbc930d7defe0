// wavefront.cuh — row-progress synchronisation of the two macroblock-wavefront kernels (k_intra, k_deblock).
//
// A picture is cut into BANDS of WF_ROWS consecutive MB rows (pair rows under MBAFF); one CTA owns one band
// and each of its warps owns one row.  Row r may work on column x once row r-1 has published progress
// >= min(x+2, width) (left, top, top-left and top-right neighbours final).
//   * rows inside a band hand over through SHARED-memory counters (fence.cta + volatile store: tens of cycles);
//   * only the first row of a band waits on a GLOBAL counter written (st.release.gpu) by the last row of the
//     band above, which lives in another CTA, possibly on another SM.
// CTAs take (band, picture) tickets from an atomic counter, picture-interleaved and band-major, so the CTA a
// band waits for always holds a smaller ticket and is already resident or finished: no deadlock however the
// hardware schedules CTAs.  Samples always travel through global memory (L2).
#pragma once
#include <cuda_runtime.h>

#ifndef WF_ROWS
#define WF_ROWS 16                       // rows (warps) per band CTA
#endif
#define WF_THREADS (WF_ROWS * 32)
#ifndef WF_POLL_NS
#define WF_POLL_NS 100                  // waiting warps sleep between polls so they do not steal issue slots from working warps
#endif

__device__ __forceinline__ int ld_relaxed_flag(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Waiting inside a band: the waiting warp parks on an mbarrier (one per row, completed by every publish of that row) instead
// of polling the shared counter — ncu showed two thirds of k_intra's executed instructions in the nanosleep poll loop.  The
// counter stays the source of truth: the waiter re-reads it after every wake-up (mbarrier.try_wait also returns after a
// hardware time limit, so a stale phase guess costs a delay, never a hang).
#ifndef WF_PARK_NS
#define WF_PARK_NS 1000
#endif
#ifndef WF_MBAR
#define WF_MBAR 1
#endif
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" :: "r"(bar) : "memory");
}
__device__ __forceinline__ int mbar_try_wait(uint32_t bar, int parity) {
    int ok;
    // suspend-time hint (WF_PARK_NS > 0).  Measured on B200 (round 2): without a hint the wait comes back after ~70 ns whether or not the
    // phase completed, so the loop around it polls — ncu counted as many warp instructions in that loop as in the filters (k_deblock3) and
    // 42 M loop trips per k_intra launch.  With a hint the warp stays parked until the arrival or the time limit; kernel times are the same
    // within noise (175.8 vs 175.7 ms per step), the issue slots go to the look-ahead stream's kernels instead of the poll.
    if (WF_PARK_NS > 0)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.s32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity), "r"(WF_PARK_NS) : "memory");
    else
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.s32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

struct RowSync {
    uint32_t bar_above, bar_mine;   // mbarriers (shared addresses) of the row above / of this row
    int gen;                        // arrivals this row has made on bar_mine so far
    volatile int *s_above;   // shared progress word of the row above (same band), or nullptr
    volatile int *s_mine;    // shared progress word of this row: (announced arrivals << 16) | progress
    int *g_above;            // global counter of the row above (other band), or nullptr when this is row 0
    int *g_mine;             // global counter of this row, or nullptr when no other band reads it
    int seen;                // last observed progress of the row above
    int width;
};

__device__ __forceinline__ RowSync rs_init(int *s_prog, uint64_t *s_bar, int warp, int row, int rows, int *g_prog, int width, int band_rows = WF_ROWS) {
    RowSync r;
    r.bar_mine = smem_addr(s_bar + warp); r.bar_above = warp > 0 ? smem_addr(s_bar + warp - 1) : 0; r.gen = 0;
    r.s_mine = s_prog + warp;
    r.s_above = warp > 0 ? s_prog + warp - 1 : nullptr;
    r.g_above = (warp == 0 && row > 0) ? g_prog + row - 1 : nullptr;
    r.g_mine = (warp == band_rows - 1 && row + 1 < rows) ? g_prog + row : nullptr;
    r.seen = row == 0 ? width : 0;
    r.width = width;
    return r;
}

// Shared hand-over protocol (lane 0 of the publishing row): write (gen + 1) << 16 | progress, THEN arrive (which completes phase
// number gen of the row's mbarrier).  A waiter that reads (g, s) with s too small parks on phase g: if arrival g - 1 has been
// announced but not yet performed, the wait returns at once and the word is read again; once phase g is current the warp
// sleeps in hardware until the next publish.  No wake-up can be lost and the waiter keeps no count of its own.
__device__ __forceinline__ void rs_store_mine(RowSync &r, int v) {
    r.gen++;
    *r.s_mine = (r.gen << 16) | v;
    if (WF_MBAR) mbar_arrive(r.bar_mine);
    if (r.g_mine) st_release_gpu(r.g_mine, v);
}

// publish "columns < v of this row are final".  Call with the whole warp converged.
__device__ __forceinline__ void rs_publish(RowSync &r, int v, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) rs_store_mine(r, v);
}

// block until the row above has published >= need; `mine` = what this row can publish meanwhile
__device__ __forceinline__ void rs_wait(RowSync &r, int need, int mine, int lane) {
    if (r.seen >= need) return;
    if (lane == 0) {
        rs_store_mine(r, mine);                    // nothing of ours is pending: everything left of `mine` was published with a fence
        int s;
        if (r.s_above) {
            for (;;) {
                const int w = *r.s_above;
                s = w & 0xffff;
                if (s >= need) break;
                if (WF_MBAR) mbar_try_wait(r.bar_above, (w >> 16) & 1); else __nanosleep(WF_POLL_NS);
            }
        } else {
            // another band (another CTA): poll the global counter, sleeping longer the further away the row above still is
            while ((s = ld_relaxed_flag(r.g_above)) < need) __nanosleep(need - s > 2 ? 8 * WF_POLL_NS : WF_POLL_NS);
            __threadfence();
        }
        r.seen = s;
    }
    r.seen = __shfl_sync(0xffffffffu, r.seen, 0);
    __threadfence_block();
    __syncwarp();
}

// non-blocking variant: true when the row above has already published >= need
__device__ __forceinline__ bool rs_try(RowSync &r, int need, int lane) {
    if (r.seen >= need) return true;
    int s = 0;
    if (lane == 0) s = r.s_above ? (*r.s_above & 0xffff) : ld_relaxed_flag(r.g_above);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s < need) return false;
    r.seen = s;
    if (r.s_above) __threadfence_block(); else __threadfence();
    __syncwarp();
    return true;
}
