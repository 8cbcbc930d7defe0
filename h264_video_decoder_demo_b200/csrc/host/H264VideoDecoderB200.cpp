// H264VideoDecoderB200.cpp — see include/H264VideoDecoderB200.h.  Plain C++ over the C ABI of the CUDA engine
// (h264_recon_b200.h); no CUDA headers, no oracle, no CPU reconstruction.
#include "H264VideoDecoderB200.h"
#include "h264_recon_b200.h"
#include "h264_front_b200.h"
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {
#pragma pack(push, 1)
struct FileHdr { char magic[8]; uint32_t version, width_mbs, height_mbs, n_pics, n_out, hdr_bytes, pichdr_bytes, reserved; };
struct PicHdr {
    int32_t decode_idx, dst_surface, clear_surface, has_inter, deblock_enable, deblock_stop_mb;
    int32_t mbaff, cqp0, cqp1, n_weights, custom_scaling, slice_type, poc, n_na;
    uint32_t n_coefs, nal_ref_idc;
    uint64_t sum_pre, sum_post;
};
struct OutRec { int32_t decode_idx, pad; uint64_t sum; };
#pragma pack(pop)
struct Pic { PicHdr h; H264B2PicParams p; };
}

CH264VideoDecoderB200::CH264VideoDecoderB200() : m_output_frame_callback(nullptr), m_userData(nullptr), m_device(0) { m_error[0] = 0; }
CH264VideoDecoderB200::~CH264VideoDecoderB200() { unInit(); }
int CH264VideoDecoderB200::init() { return 0; }
int CH264VideoDecoderB200::unInit() { return 0; }
int CH264VideoDecoderB200::set_output_frame_callback_functuin(output_frame_callback_b200 cb, void *userData) { m_output_frame_callback = cb; m_userData = userData; return 0; }
int CH264VideoDecoderB200::set_device(int device) { m_device = device; return 0; }

#define FAIL(code, ...) do { snprintf(m_error, sizeof m_error, __VA_ARGS__); ret = (code); goto done; } while (0)

int CH264VideoDecoderB200::open(const char *url) {
    int ret = 0;
    H264B2Context *ctx = nullptr;
    uint8_t *blob = nullptr, *frame = nullptr;
    FILE *f = nullptr;
    std::vector<Pic> pics;
    std::vector<OutRec> outs;
    std::vector<char> decoded;
    size_t next_out = 0;
    FileHdr fh;
    if (!url) { snprintf(m_error, sizeof m_error, "open: null url"); return -1; }
    f = fopen(url, "rb");
    if (!f) FAIL(-1, "open: cannot open %s", url);
    if (fread(&fh, sizeof fh, 1, f) != 1 || memcmp(fh.magic, "H264B2RP", 8)) {
        // not a pre-parsed container: an Annex-B byte stream, like the reference's open() takes (H264VideoDecoder.cpp:48)
        fclose(f); f = nullptr;
        return open_bitstream(url);
    }
    if (fh.version != 1 || fh.hdr_bytes != sizeof fh || fh.pichdr_bytes != sizeof(PicHdr))
        FAIL(-1, "open: %s is not a picture container", url);
    {
        if (fseek(f, 0, SEEK_END)) FAIL(-1, "open: seek failed");
        const long total = ftell(f);
        if (h264b2_create(&ctx, m_device, 1, 17, (int)fh.width_mbs, (int)fh.height_mbs)) FAIL(-3, "open: %s", h264b2_last_error());
        // the whole container goes into page-locked memory: h264b2_submit DMAs each picture's arrays from it
        if (h264b2_host_alloc(ctx, (size_t)total, (void **)&blob)) FAIL(-3, "open: %s", h264b2_last_error());
        fseek(f, 0, SEEK_SET);
        if (fread(blob, 1, (size_t)total, f) != (size_t)total) FAIL(-1, "open: short read");
        const size_t nmb = (size_t)fh.width_mbs * fh.height_mbs;
        size_t off = sizeof fh;
        for (uint32_t i = 0; i < fh.n_pics; i++) {
            Pic pc; memset(&pc, 0, sizeof pc);
            if (off + sizeof(PicHdr) > (size_t)total) FAIL(-1, "open: truncated container");
            memcpy(&pc.h, blob + off, sizeof(PicHdr)); off += sizeof(PicHdr);
            if (pc.h.n_weights < 1 || pc.h.n_weights > 65536 || (size_t)pc.h.n_coefs > (size_t)total / 2) FAIL(-1, "open: corrupt picture header %u", i);
            H264B2PicParams &p = pc.p;
            p.width_mbs = (int)fh.width_mbs; p.height_mbs = (int)fh.height_mbs; p.mbaff_frame_flag = pc.h.mbaff;
            p.chroma_qp_offset[0] = pc.h.cqp0; p.chroma_qp_offset[1] = pc.h.cqp1;
            p.dst_surface = pc.h.dst_surface; p.clear_surface = pc.h.clear_surface; p.has_inter = pc.h.has_inter;
            p.deblock_enable = pc.h.deblock_enable; p.deblock_stop_mb = pc.h.deblock_stop_mb;
            p.n_weights = pc.h.n_weights; p.n_coefs = pc.h.n_coefs; p.custom_scaling = pc.h.custom_scaling;
            p.mb_info = (const H264B2MbInfo *)(blob + off); off += nmb * sizeof(H264B2MbInfo);
            p.intra_modes = (const uint64_t *)(blob + off); off += nmb * 8;
            p.coef_offset = (const uint32_t *)(blob + off); off += nmb * 4;
            if (pc.h.has_inter) { p.motion = (const H264B2MbMotion *)(blob + off); off += nmb * sizeof(H264B2MbMotion); }
            p.weights = (const H264B2Weight *)(blob + off); off += (size_t)pc.h.n_weights * sizeof(H264B2Weight);
            p.coefs = (const int16_t *)(blob + off); off += (size_t)pc.h.n_coefs * 2;
            if (pc.h.custom_scaling) { p.level_scale4 = (const int16_t *)(blob + off); off += 2 * 2 * 6 * 16 * 2; p.level_scale8 = (const int16_t *)(blob + off); off += 2 * 2 * 6 * 64 * 2; }
            if (off > (size_t)total) FAIL(-1, "open: truncated container");
            pics.push_back(pc);
        }
        if (off + (size_t)fh.n_out * sizeof(OutRec) > (size_t)total) FAIL(-1, "open: truncated container");
        outs.resize(fh.n_out);
        if (fh.n_out) memcpy(outs.data(), blob + off, fh.n_out * sizeof(OutRec));
        const size_t frame_bytes = nmb * 384;
        if (h264b2_host_alloc(ctx, frame_bytes, (void **)&frame)) FAIL(-3, "open: %s", h264b2_last_error());
        decoded.assign(pics.size(), 0);
        const int W = (int)fh.width_mbs * 16, H = (int)fh.height_mbs * 16;
        int stop = 0;
        // Decode in decoding order.  A picture is handed to the callback as soon as it and every picture before
        // it in OUTPUT order have been reconstructed: never later than the reference's bumping process
        // (H264PicturesGOP.cpp:89-165) would, so its surface has not been recycled yet; the order is the reference's.
        for (size_t i = 0; i < pics.size() && !stop; i++) {
            const int32_t sid = 0;
            if (h264b2_submit(ctx, 1, &sid, &pics[i].p)) FAIL(-3, "open: %s", h264b2_last_error());
            decoded[i] = 1;
            while (next_out < outs.size() && !stop) {
                const int di = outs[next_out].decode_idx;
                if (di < 0 || (size_t)di >= pics.size() || !decoded[di]) break;
                if (h264b2_read_picture(ctx, 0, pics[di].h.dst_surface, frame)) FAIL(-3, "open: %s", h264b2_last_error());
                CH264PictureB200 out; memset(&out, 0, sizeof out);
                CH264PictureBaseB200 &b = out.m_picture_frame;
                b.m_pic_buff_luma = frame; b.m_pic_buff_cb = frame + (size_t)W * H; b.m_pic_buff_cr = b.m_pic_buff_cb + (size_t)(W / 2) * (H / 2);
                b.PicWidthInSamplesL = W; b.PicHeightInSamplesL = H; b.PicWidthInSamplesC = W / 2; b.PicHeightInSamplesC = H / 2;
                b.PicOrderCnt = pics[di].h.poc; b.m_PicNumCnt = di; b.slice_type = pics[di].h.slice_type; b.MbaffFrameFlag = pics[di].h.mbaff;
                next_out++;
                if (m_output_frame_callback && m_output_frame_callback(&out, m_userData, H264_DECODE_ERROR_CODE_NO) != 0) stop = 1;   // VD:119-124
            }
        }
        if (!stop && m_output_frame_callback) m_output_frame_callback(nullptr, m_userData, H264_DECODE_ERROR_CODE_FILE_END);          // VD:372-374
    }
done:
    if (f) fclose(f);
    if (ctx) { if (frame) h264b2_host_free(ctx, frame); if (blob) h264b2_host_free(ctx, blob); h264b2_destroy(ctx); }
    return ret;
}

// ---- Annex-B path: host entropy/derivation stage (h264_front_b200.h) on a producer thread, GPU reconstruction and the
// output callback on the calling thread (the reference invokes the callback synchronously on the caller, VD:396-432).
namespace {
struct PinCtx { H264B2Context *ctx; };
void *pin_alloc(void *u, size_t n) { void *p = nullptr; return h264b2_host_alloc(((PinCtx *)u)->ctx, n, &p) == 0 ? p : nullptr; }
void pin_free(void *u, void *p) { h264b2_host_free(((PinCtx *)u)->ctx, p); }
struct EvQueue {
    std::mutex mu; std::condition_variable cv; std::deque<H264B2FrontEvent> q; bool done = false, cancel = false; int err = 0;
    void push(const H264B2FrontEvent &e) { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return q.size() < 12 || cancel; }); q.push_back(e); cv.notify_all(); }
    bool pop(H264B2FrontEvent *e) { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return !q.empty() || done; }); if (q.empty()) return false; *e = q.front(); q.pop_front(); cv.notify_all(); return true; }
};
}

int CH264VideoDecoderB200::open_bitstream(const char *url) {
    int ret = 0;
    H264B2Context *ctx = nullptr;
    H264B2Front *fe = nullptr, *probe = nullptr;
    uint8_t *frame = nullptr;
    PinCtx pc = {nullptr};
    EvQueue Q;
    std::thread producer;
    std::deque<void *> inflight;       // picture blocks whose DMA may still be pending (h264b2_submit keeps 3 batches in flight)
    int wmb = 0, hmb = 0, n_frames = 0;
    H264B2FrontEvent ev;
    H264B2StreamInfo sinfo; memset(&sinfo, 0, sizeof sinfo);
    // picture size: parse up to the first picture with a throw-away front end (the context needs the size up front)
    if (h264b2_front_create(&probe, nullptr, nullptr, nullptr) || h264b2_front_open_file(probe, url)) FAIL(-1, "open: %s", probe ? h264b2_front_last_error(probe) : "out of memory");
    for (;;) {
        if (h264b2_front_next(probe, &ev) < 0) FAIL(-1, "open: %s", h264b2_front_last_error(probe));
        if (ev.kind == H264B2_EV_PICTURE) { wmb = ev.width_mbs; hmb = ev.height_mbs; h264b2_front_stream_info(probe, &sinfo); break; }
        if (ev.kind == H264B2_EV_END) FAIL(-1, "open: %s holds no decodable picture", url);
    }
    h264b2_front_destroy(probe); probe = nullptr;
    if (h264b2_create(&ctx, m_device, 1, 17, wmb, hmb)) FAIL(-3, "open: %s", h264b2_last_error());
    pc.ctx = ctx;
    if (h264b2_host_alloc(ctx, (size_t)wmb * hmb * 384, (void **)&frame)) FAIL(-3, "open: %s", h264b2_last_error());
    if (h264b2_front_create(&fe, pin_alloc, pin_free, &pc) || h264b2_front_open_file(fe, url)) FAIL(-1, "open: %s", fe ? h264b2_front_last_error(fe) : "out of memory");
    h264b2_front_set_packed(fe, getenv("H264B2_PACKED_ARRAYS") ? (H264B2_PACKED_COEFS | H264B2_PACKED_MOTION) : 0);      // H264B2_PACKED_ARRAYS=1: packed levels and motion over PCIe. Off by default: this path is bound by the host parser (~70 pictures/s/thread, < 1 GB/s of PCIe), and packing costs it 1.5 ms per picture; it pays when submits run at > 10k pictures/s (bench e2e)
    producer = std::thread([&] {
        for (;;) {
            H264B2FrontEvent e;
            const int r = h264b2_front_next(fe, &e);
            if (r < 0) { std::lock_guard<std::mutex> l(Q.mu); Q.err = r; Q.done = true; Q.cv.notify_all(); return; }
            Q.push(e);
            if (e.kind == H264B2_EV_END || Q.cancel) { std::lock_guard<std::mutex> l(Q.mu); Q.done = true; Q.cv.notify_all(); return; }
        }
    });
    {
        const int W = wmb * 16, H = hmb * 16;
        std::vector<int> surf_poc(17, 0), surf_idx(17, 0), surf_type(17, 0), surf_mbaff(17, 0);
        bool ended = false;
        while (Q.pop(&ev)) {
            if (ev.kind == H264B2_EV_PICTURE) {
                const int32_t sid = 0;
                if (!ev.block) { ret = -3; snprintf(m_error, sizeof m_error, "open: out of page-locked memory"); break; }
                if (h264b2_submit(ctx, 1, &sid, &ev.params)) { ret = -3; snprintf(m_error, sizeof m_error, "open: %s", h264b2_last_error()); break; }
                inflight.push_back(ev.block);
                while (inflight.size() > H264B2_SUBMIT_DEPTH) { h264b2_front_release(fe, inflight.front()); inflight.pop_front(); }
                surf_poc[ev.surface] = ev.hdr.poc; surf_idx[ev.surface] = ev.decode_idx; surf_type[ev.surface] = ev.hdr.slice_type; surf_mbaff[ev.surface] = ev.hdr.mbaff;
            } else if (ev.kind == H264B2_EV_OUTPUT) {
                if (h264b2_read_picture(ctx, 0, ev.surface, frame)) { ret = -3; snprintf(m_error, sizeof m_error, "open: %s", h264b2_last_error()); break; }
                CH264PictureB200 out; memset(&out, 0, sizeof out);
                CH264PictureBaseB200 &b = out.m_picture_frame;
                b.m_pic_buff_luma = frame; b.m_pic_buff_cb = frame + (size_t)W * H; b.m_pic_buff_cr = b.m_pic_buff_cb + (size_t)(W / 2) * (H / 2);
                b.PicWidthInSamplesL = W; b.PicHeightInSamplesL = H; b.PicWidthInSamplesC = W / 2; b.PicHeightInSamplesC = H / 2;
                b.PicOrderCnt = surf_poc[ev.surface]; b.m_PicNumCnt = surf_idx[ev.surface]; b.slice_type = surf_type[ev.surface]; b.MbaffFrameFlag = surf_mbaff[ev.surface];
                b.profile_idc = sinfo.profile_idc; b.level_idc = sinfo.level_idc; b.entropy_coding_mode_flag = sinfo.entropy_coding_mode_flag; b.fps = (float)sinfo.fps;
                n_frames++;
                if (m_output_frame_callback && m_output_frame_callback(&out, m_userData, H264_DECODE_ERROR_CODE_NO) != 0) break;      // VD:119-124: stop
            } else { ended = true; break; }
        }
        { std::lock_guard<std::mutex> l(Q.mu); Q.cancel = true; Q.cv.notify_all(); }
        // drain so that the producer can finish
        while (Q.pop(&ev)) { if (ev.kind == H264B2_EV_PICTURE && ev.block) h264b2_front_release(fe, ev.block); }
        if (producer.joinable()) producer.join();
        if (!ret && Q.err) { ret = -1; snprintf(m_error, sizeof m_error, "open: %s", h264b2_front_last_error(fe)); }
        if (!ret && ended && m_output_frame_callback) m_output_frame_callback(nullptr, m_userData, H264_DECODE_ERROR_CODE_FILE_END);   // VD:372-374
    }
done:
    if (producer.joinable()) { { std::lock_guard<std::mutex> l(Q.mu); Q.cancel = true; Q.cv.notify_all(); } while (Q.pop(&ev)) {} producer.join(); }
    if (probe) h264b2_front_destroy(probe);
    if (ctx) h264b2_sync(ctx);
    if (fe) h264b2_front_destroy(fe);
    if (ctx) { if (frame) h264b2_host_free(ctx, frame); h264b2_destroy(ctx); }
    (void)n_frames;
    return ret;
}

extern "C" {
void *h264b2_decoder_create(void) { return new CH264VideoDecoderB200(); }
void h264b2_decoder_destroy(void *d) { delete (CH264VideoDecoderB200 *)d; }
int h264b2_decoder_set_callback(void *d, output_frame_callback_b200 cb, void *u) { return d ? ((CH264VideoDecoderB200 *)d)->set_output_frame_callback_functuin(cb, u) : -1; }
int h264b2_decoder_set_device(void *d, int device) { return d ? ((CH264VideoDecoderB200 *)d)->set_device(device) : -1; }
int h264b2_decoder_open(void *d, const char *url) { return d ? ((CH264VideoDecoderB200 *)d)->open(url) : -1; }
const char *h264b2_decoder_last_error(void *d) { return d ? ((CH264VideoDecoderB200 *)d)->last_error() : "null decoder"; }
}
