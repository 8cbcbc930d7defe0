// h264_multi.cpp — see include/h264_multi_b200.h.  Parser thread pool -> per-stream event queues -> one submit thread.
#include <stdlib.h>
#include "h264_multi_b200.h"
#include "h264_front_b200.h"
#include "h264_recon_b200.h"
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <stdio.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>

namespace {
struct Pin { H264B2Context *ctx; };
void *pin_alloc(void *u, size_t n) { void *p = nullptr; return h264b2_host_alloc(((Pin *)u)->ctx, n, &p) == 0 ? p : nullptr; }
void pin_free(void *u, void *p) { h264b2_host_free(((Pin *)u)->ctx, p); }

struct Stream {                 // one decoding unit: a whole stream, or one closed-GOP shard of it
    int owner = 0; size_t begin = 0, end = 0; int more_follows = 0;
    std::vector<uint64_t> sums;
    H264B2Front *fe = nullptr;
    std::deque<H264B2FrontEvent> q;
    int pics_queued = 0;
    bool parsed_all = false, finished = false, claimed = false;      // claimed: a parser thread is inside this stream's front end
    uint64_t hash = 0;
    int ring = 0;
};
enum { RING = 4 };
}

extern "C" int h264b2_multi_decode(int device, int n_inputs, const char *const *paths, int n_threads, int flags,
                                   uint64_t *stream_hash, H264B2MultiStats *stats, char *err, size_t err_len) {
    auto fail = [&](int code, const std::string &m) { if (err && err_len) snprintf(err, err_len, "%s", m.c_str()); return code; };
    if (n_inputs <= 0 || !paths || n_threads <= 0) return fail(-1, "bad argument");
    const int readback = flags & H264B2_MULTI_READBACK, split = flags & H264B2_MULTI_SPLIT_GOPS;
    // the byte streams (identical paths share one buffer), and the decoding units: whole streams or closed-GOP shards
    std::vector<std::vector<uint8_t>> bufs; std::vector<std::string> names; std::vector<int> buf_of(n_inputs);
    for (int i = 0; i < n_inputs; i++) {
        int b = -1;
        for (size_t k = 0; k < names.size(); k++) if (names[k] == paths[i]) b = (int)k;
        if (b < 0) {
            FILE *fp = fopen(paths[i], "rb");
            if (!fp) return fail(-1, std::string("cannot open ") + paths[i]);
            fseek(fp, 0, SEEK_END); const long n = ftell(fp); fseek(fp, 0, SEEK_SET);
            std::vector<uint8_t> v((size_t)n + 16, 0);
            if (n > 0 && fread(v.data(), 1, (size_t)n, fp) != (size_t)n) { fclose(fp); return fail(-1, std::string("short read: ") + paths[i]); }
            fclose(fp);
            v.resize((size_t)n);
            bufs.push_back(std::move(v)); names.push_back(paths[i]); b = (int)bufs.size() - 1;
        }
        buf_of[i] = b;
    }
    std::vector<Stream> st;
    for (int i = 0; i < n_inputs; i++) {
        const std::vector<uint8_t> &v = bufs[buf_of[i]];
        std::vector<size_t> offs(1, 0);
        if (split) { const int n = h264b2_front_gop_offsets(v.data(), v.size(), nullptr, 0); if (n > 1) { offs.assign((size_t)n, 0); h264b2_front_gop_offsets(v.data(), v.size(), offs.data(), n); } }
        for (size_t g = 0; g < offs.size(); g++) {
            Stream x; x.owner = i; x.begin = offs[g]; x.end = g + 1 < offs.size() ? offs[g + 1] : v.size(); x.more_follows = g + 1 < offs.size();
            st.push_back(std::move(x));
        }
    }
    const int n_streams = (int)st.size();
    // pictures a stream's parser may run ahead of the submit thread (H264B2_MULTI_QUEUE_DEPTH, default 3; a queued picture holds one page-locked block)
    int QUEUE_DEPTH = 3;
    if (const char *e = getenv("H264B2_MULTI_QUEUE_DEPTH")) { const int v = atoi(e); if (v >= 1 && v <= 64) QUEUE_DEPTH = v; }
    if (n_threads > n_streams) n_threads = n_streams;
    // picture size from the first stream (all streams of one context share it)
    int wmb = 0, hmb = 0;
    {
        H264B2Front *p = nullptr; H264B2FrontEvent ev;
        if (h264b2_front_create(&p, nullptr, nullptr, nullptr) || h264b2_front_open_file(p, paths[0])) { std::string m = p ? h264b2_front_last_error(p) : "out of memory"; if (p) h264b2_front_destroy(p); return fail(-1, m); }
        for (;;) {
            if (h264b2_front_next(p, &ev) < 0) { std::string m = h264b2_front_last_error(p); h264b2_front_destroy(p); return fail(-1, m); }
            if (ev.kind == H264B2_EV_PICTURE) { wmb = ev.width_mbs; hmb = ev.height_mbs; break; }
            if (ev.kind == H264B2_EV_END) { h264b2_front_destroy(p); return fail(-1, std::string(paths[0]) + " holds no picture"); }
        }
        h264b2_front_destroy(p);
    }
    H264B2Context *ctx = nullptr;
    if (h264b2_create(&ctx, device, n_streams, 17, wmb, hmb)) return fail(-3, h264b2_last_error());
    Pin pin = {ctx};
    const size_t frame_bytes = (size_t)wmb * hmb * 384;
    uint8_t *frames = nullptr;
    if (readback && h264b2_host_alloc(ctx, frame_bytes * n_streams * RING, (void **)&frames)) { std::string m = h264b2_last_error(); h264b2_destroy(ctx); return fail(-3, m); }
    std::string first_error;
    for (int s = 0; s < n_streams; s++) {
        const std::vector<uint8_t> &v = bufs[buf_of[st[s].owner]];
        if (h264b2_front_create(&st[s].fe, pin_alloc, pin_free, &pin) || h264b2_front_open_range(st[s].fe, v.data(), v.size(), st[s].begin, st[s].end, st[s].more_follows)) { first_error = st[s].fe ? h264b2_front_last_error(st[s].fe) : "out of memory"; break; }
        h264b2_front_set_packed(st[s].fe, getenv("H264B2_PACKED_ARRAYS") ? (H264B2_PACKED_COEFS | H264B2_PACKED_MOTION) : 0);      // H264B2_PACKED_ARRAYS=1: packed levels and motion over PCIe. Off by default: this path is bound by the host parser (~70 pictures/s/thread, < 1 GB/s of PCIe), and packing costs it 1.5 ms per picture; it pays when submits run at > 10k pictures/s (bench e2e)
    }
    H264B2MultiStats S; memset(&S, 0, sizeof S);
    S.threads = n_threads; S.streams = n_inputs; S.units = n_streams; S.width_mbs = wmb; S.height_mbs = hmb;
    int rc = 0;
    if (first_error.empty()) {
        std::mutex mu; std::condition_variable cv_space, cv_data;
        bool cancel = false;
        std::vector<double> busy(n_threads, 0.0);
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        // Parser threads claim streams dynamically: whichever unclaimed stream has the emptiest queue (so the submit thread's batches stay
        // wide and no thread idles while another still has several streams to go — the first version dealt streams s -> thread s mod T,
        // which left threads idle whenever T did not divide the number of streams or the streams differed in length).  One claim = one
        // front-end event; a front end is only ever inside one thread at a time and changes hands under `mu`.
        for (int t = 0; t < n_threads; t++)
            pool.emplace_back([&, t] {
                int next = t % n_streams;
                for (;;) {
                    int s = -1;
                    {
                        std::unique_lock<std::mutex> l(mu);
                        for (;;) {
                            if (cancel) return;
                            bool all_done = true; int best_q = QUEUE_DEPTH;
                            for (int i = 0, c = next; i < n_streams; i++, c = c + 1 == n_streams ? 0 : c + 1) {
                                const Stream &x = st[c];
                                if (x.parsed_all) continue;
                                all_done = false;
                                if (!x.claimed && x.pics_queued < best_q) { s = c; best_q = x.pics_queued; }
                            }
                            if (all_done) return;
                            if (s >= 0) { st[s].claimed = true; break; }
                            cv_space.wait_for(l, std::chrono::milliseconds(2));
                        }
                    }
                    Stream &x = st[s];
                    const auto b0 = std::chrono::steady_clock::now();
                    H264B2FrontEvent ev;
                    const int r = h264b2_front_next(x.fe, &ev);
                    busy[t] += std::chrono::duration<double>(std::chrono::steady_clock::now() - b0).count();
                    std::lock_guard<std::mutex> l(mu);
                    if (r < 0) { if (first_error.empty()) first_error = std::string(paths[x.owner]) + ": " + h264b2_front_last_error(x.fe); memset(&ev, 0, sizeof ev); ev.kind = H264B2_EV_END; }
                    x.q.push_back(ev);
                    if (ev.kind == H264B2_EV_PICTURE) x.pics_queued++;
                    if (ev.kind == H264B2_EV_END) x.parsed_all = true;
                    x.claimed = false;
                    cv_data.notify_one();
                    next = s + 1 == n_streams ? 0 : s + 1;
                }
            });
        // ---- submit thread (this one)
        std::deque<std::vector<std::pair<int, void *>>> inflight;       // blocks of the last submits (DMA may be pending for H264B2_SUBMIT_DEPTH of them)
        int active = n_streams, rounds_since_sync = 0;
        std::vector<int32_t> sids, surfs; std::vector<uint8_t *> hptr; std::vector<H264B2PicParams> pics; std::vector<uint64_t> sums;
        while (active > 0 && rc == 0) {
            std::vector<std::vector<std::pair<int, int>>> out_rounds;  // [round] -> (stream, surface)
            std::vector<std::pair<int, H264B2FrontEvent>> batch;
            {
                std::unique_lock<std::mutex> l(mu);
                cv_data.wait(l, [&] { for (auto &x : st) if (!x.finished && !x.q.empty()) return true; return false; });
                for (int s = 0; s < n_streams; s++) {
                    Stream &x = st[s];
                    size_t r = 0;
                    while (!x.q.empty() && x.q.front().kind == H264B2_EV_OUTPUT) {
                        if (out_rounds.size() <= r) out_rounds.emplace_back();
                        out_rounds[r++].push_back({s, x.q.front().surface});
                        x.q.pop_front();
                    }
                    if (!x.q.empty() && x.q.front().kind == H264B2_EV_PICTURE) { batch.push_back({s, x.q.front()}); x.q.pop_front(); x.pics_queued--; }
                    else if (!x.q.empty() && x.q.front().kind == H264B2_EV_END) { x.q.pop_front(); x.finished = true; active--; }
                }
                cv_space.notify_all();
            }
            for (auto &round : out_rounds) {
                sids.clear(); surfs.clear(); hptr.clear();
                for (auto &o : round) { sids.push_back(o.first); surfs.push_back(o.second); }
                if (stream_hash) {
                    sums.assign(round.size(), 0);
                    if (h264b2_checksum_pictures(ctx, (int)round.size(), sids.data(), surfs.data(), sums.data())) { rc = -3; first_error = h264b2_last_error(); break; }
                    for (size_t i = 0; i < round.size(); i++) st[round[i].first].sums.push_back(sums[i]);
                }
                if (readback) {
                    bool wrap = false;
                    for (auto &o : round) { Stream &x = st[o.first]; hptr.push_back(frames + ((size_t)o.first * RING + x.ring) * frame_bytes); x.ring = (x.ring + 1) % RING; if (x.ring == 0) wrap = true; }
                    if (wrap || ++rounds_since_sync >= RING - 1) { h264b2_sync(ctx); rounds_since_sync = 0; }     // host ring slots are reused only after their copies completed
                    if (h264b2_read_pictures_async(ctx, (int)round.size(), sids.data(), surfs.data(), hptr.data())) { rc = -3; first_error = h264b2_last_error(); break; }
                    S.d2h_bytes += (int64_t)round.size() * frame_bytes;
                }
                S.frames_out += (int64_t)round.size();
            }
            if (rc) break;
            if (!batch.empty()) {
                sids.clear(); pics.clear();
                std::vector<std::pair<int, void *>> blocks;
                for (auto &b : batch) {
                    if (!b.second.block) { rc = -3; first_error = "out of page-locked memory"; break; }
                    sids.push_back(b.first); pics.push_back(b.second.params); blocks.push_back({b.first, b.second.block}); S.h2d_bytes += (int64_t)b.second.block_bytes;
                }
                if (rc) break;
                if (h264b2_submit(ctx, (int)pics.size(), sids.data(), pics.data())) { rc = -3; first_error = h264b2_last_error(); break; }
                S.pictures += (int64_t)pics.size(); S.submits++;
                inflight.push_back(blocks);
                while (inflight.size() > H264B2_SUBMIT_DEPTH) { for (auto &b : inflight.front()) h264b2_front_release(st[b.first].fe, b.second); inflight.pop_front(); }
            }
        }
        { std::lock_guard<std::mutex> l(mu); cancel = true; cv_space.notify_all(); }
        for (auto &t : pool) t.join();
        h264b2_sync(ctx);
        S.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (double b : busy) S.parse_seconds += b;
        if (!first_error.empty() && rc == 0) rc = -1;
    } else rc = -1;
    if (stream_hash) {          // per input stream: chain over its output frames, GOP shards in stream order
        for (int i = 0; i < n_inputs; i++) stream_hash[i] = 0;
        for (auto &x : st) for (uint64_t v : x.sums) stream_hash[x.owner] = stream_hash[x.owner] * 0x100000001B3ULL + v;
    }
    for (auto &x : st) if (x.fe) h264b2_front_destroy(x.fe);
    if (frames) h264b2_host_free(ctx, frames);
    h264b2_destroy(ctx);
    if (stats) *stats = S;
    if (rc) return fail(rc, first_error);
    return 0;
}
