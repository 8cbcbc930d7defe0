// H264VideoDecoderB200.cpp — see include/H264VideoDecoderB200.h.  Plain C++ over the C ABI of the CUDA engine
// (h264_recon_b200.h); no CUDA headers, no oracle, no CPU reconstruction.
#include "H264VideoDecoderB200.h"
#include "h264_recon_b200.h"
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
    {
        const size_t n = strlen(url);
        if (n > 5 && (!strcmp(url + n - 5, ".h264") || !strcmp(url + n - 4, ".264")))
            FAIL(-2, "open: %s is a raw byte stream; the native entropy front end is not part of this build (feed a pre-parsed picture container); no CPU fallback", url);
    }
    f = fopen(url, "rb");
    if (!f) FAIL(-1, "open: cannot open %s", url);
    if (fread(&fh, sizeof fh, 1, f) != 1 || memcmp(fh.magic, "H264B2RP", 8) || fh.version != 1 || fh.hdr_bytes != sizeof fh || fh.pichdr_bytes != sizeof(PicHdr))
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

extern "C" {
void *h264b2_decoder_create(void) { return new CH264VideoDecoderB200(); }
void h264b2_decoder_destroy(void *d) { delete (CH264VideoDecoderB200 *)d; }
int h264b2_decoder_set_callback(void *d, output_frame_callback_b200 cb, void *u) { return d ? ((CH264VideoDecoderB200 *)d)->set_output_frame_callback_functuin(cb, u) : -1; }
int h264b2_decoder_set_device(void *d, int device) { return d ? ((CH264VideoDecoderB200 *)d)->set_device(device) : -1; }
int h264b2_decoder_open(void *d, const char *url) { return d ? ((CH264VideoDecoderB200 *)d)->open(url) : -1; }
const char *h264b2_decoder_last_error(void *d) { return d ? ((CH264VideoDecoderB200 *)d)->last_error() : "null decoder"; }
}
