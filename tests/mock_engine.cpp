// mock_engine.cpp — TEST DOUBLE of the CUDA engine's C ABI (include/h264_recon_b200.h) for host-only thread-safety tests
// (tests/test_multi_pool_tsan.py): no pixel is produced, pictures read back as zeros.  It exists so that the multi-stream pipeline
// (parser pool, per-stream queues, block pool, submit thread) can run under ThreadSanitizer on a machine without a GPU.  Never linked
// into the product.  It keeps the contract that matters to the host side: arrays handed to submit are READ (as a DMA would) and must
// still be alive H264B2_SUBMIT_DEPTH submits later.
#include "h264_recon_b200.h"
#include <stdlib.h>
#include <string.h>
#include <deque>
#include <vector>
struct H264B2Context { int n_streams, wmb, hmb; std::deque<std::vector<const uint8_t *>> inflight; unsigned long long sink; };
extern "C" {
int h264b2_create(H264B2Context **ctx, int, int n_streams, int, int width_mbs, int height_mbs) { *ctx = new H264B2Context{n_streams, width_mbs, height_mbs, {}, 0}; return 0; }
int h264b2_destroy(H264B2Context *c) { delete c; return 0; }
int h264b2_host_alloc(H264B2Context *, size_t bytes, void **p) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : -3; }
int h264b2_host_free(H264B2Context *, void *p) { free(p); return 0; }
int h264b2_submit(H264B2Context *c, int n, const int32_t *sids, const H264B2PicParams *pics) {
    // touch the arrays of the submits still "in flight" (what the copy engine would do) and of this one
    std::vector<const uint8_t *> cur;
    for (int i = 0; i < n; i++) {
        if (sids[i] < 0 || sids[i] >= c->n_streams) return -1;
        const size_t nmb = (size_t)c->wmb * c->hmb;
        const uint8_t *p = (const uint8_t *)pics[i].mb_info;
        for (size_t k = 0; k < nmb * sizeof(H264B2MbInfo); k += 64) c->sink += p[k];
        if (pics[i].n_coefs) { const uint8_t *q = (const uint8_t *)pics[i].coefs; c->sink += q[0] + q[(size_t)pics[i].n_coefs * 2 - 1]; }
        cur.push_back(p);
    }
    for (auto &v : c->inflight) for (const uint8_t *p : v) c->sink += p[0];
    c->inflight.push_back(cur);
    while (c->inflight.size() > H264B2_SUBMIT_DEPTH - 1) c->inflight.pop_front();
    return 0;
}
int h264b2_sync(H264B2Context *c) { c->inflight.clear(); return 0; }
int h264b2_read_picture(H264B2Context *c, int, int, uint8_t *h) { memset(h, 0, (size_t)c->wmb * c->hmb * 384); return 0; }
int h264b2_read_pictures_async(H264B2Context *c, int n, const int32_t *, const int32_t *, uint8_t *const *h) { for (int i = 0; i < n; i++) memset(h[i], 0, (size_t)c->wmb * c->hmb * 384); return 0; }
int h264b2_checksum_pictures(H264B2Context *, int n, const int32_t *sids, const int32_t *surf, uint64_t *sums) { (void)sids; for (int i = 0; i < n; i++) sums[i] = (uint64_t)surf[i] * 2654435761u + 1; return 0; }      // a function of the output surface only: replicas must agree
const char *h264b2_last_error(void) { return "mock engine"; }
size_t h264b2_pack_coefs_bound(uint32_t n) { return (size_t)n * 4 + 64; }
int h264b2_pack_coefs(const int16_t *, uint32_t, void *, size_t, size_t *) { return -1; }
int h264b2_pack_motion(const H264B2MbMotion *, uint32_t, void *, size_t, size_t *) { return -1; }
}
