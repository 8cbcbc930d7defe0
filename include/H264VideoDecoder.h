/*
 * H264VideoDecoder.h — SAME-NAME shim of the reference's public decoder interface (jfu222/h264_video_decoder_demo,
 * H264VideoDecoder.h:22-43) over the B200 engine: code written against the reference — its main.cpp callback
 * (main.cpp:13-52), the player's frame handler (SDH264Player/SDH264Player/MyStatic.cpp:153-211) — compiles against this
 * header unchanged and links with libh264b2_host.so instead of the reference's objects.
 *
 *   class CH264VideoDecoder { init(); unInit(); set_output_frame_callback_functuin(cb, userData); open(url); do_callback(...); }
 *   typedef int (*output_frame_callback)(CH264Picture *outPicture, void *userData, int errorCode);
 *   outPicture->m_picture_frame.{m_pic_buff_luma, m_pic_buff_cb, m_pic_buff_cr, PicWidthInSamplesL, PicHeightInSamplesL,
 *       PicWidthInSamplesC, PicHeightInSamplesC, PicOrderCnt, m_PicNumCnt, m_h264_slice_header.{slice_type, MbaffFrameFlag,
 *       m_sps.{profile_idc, level_idc, fps}, m_pps.entropy_coding_mode_flag}, saveToBmpFile(), convertYuv420pToBgr24()}
 *
 * Not carried: m_picture_frame.m_mbs (the player's per-macroblock inspector copies the reference's 18 KB-per-MB state; the
 * engine's per-picture structure-of-arrays is available through h264_front_b200.h instead).
 * Same conventions as the reference: int returns (0 ok), callback on the calling thread in display order, a final call with
 * outPicture == NULL and errorCode == H264_DECODE_ERROR_CODE_FILE_END, a non-zero callback return stops decoding, the picture is
 * valid only inside the callback.
 */
#ifndef H264_VIDEO_DECODER_SHIM_H
#define H264_VIDEO_DECODER_SHIM_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "H264VideoDecoderB200.h"

/* H264CommonFunc.h:245-258 (slice_type 0..9) */
#define H264_SLIECE_TYPE_TO_STR(t) (((t) % 5) == 0 ? "P" : ((t) % 5) == 1 ? "B" : ((t) % 5) == 2 ? "I" : ((t) % 5) == 3 ? "SP" : "SI")

struct CH264SPS { int32_t profile_idc, level_idc; float fps; };
struct CH264PPS { int32_t entropy_coding_mode_flag; };
struct CH264SliceHeader { int32_t slice_type, MbaffFrameFlag; CH264SPS m_sps; CH264PPS m_pps; };

class CH264PictureBase {
public:
    uint8_t *m_pic_buff_luma, *m_pic_buff_cb, *m_pic_buff_cr;
    int32_t PicWidthInSamplesL, PicHeightInSamplesL, PicWidthInSamplesC, PicHeightInSamplesC;
    int32_t PicOrderCnt, m_PicNumCnt;
    CH264SliceHeader m_h264_slice_header;
    /* H264PictureBase.cpp:440-468: the reference's integer BT.601 conversion, bottom-up rows when isFlip (the BMP writer's order) */
    int convertYuv420pToBgr24(uint32_t width, uint32_t height, const uint8_t *yuv420p, uint8_t *bgr24, uint32_t widthBytesBgr24) const {
        const uint8_t *py = yuv420p, *pu = yuv420p + (size_t)width * height, *pv = pu + (size_t)width * height / 4;
        for (uint32_t y = 0; y < height; y++)
            for (uint32_t x = 0; x < width; x++) {
                const int Y = 1164 * ((int)py[(size_t)y * width + x] - 16), U = (int)pu[(size_t)(y >> 1) * (width >> 1) + (x >> 1)] - 128, V = (int)pv[(size_t)(y >> 1) * (width >> 1) + (x >> 1)] - 128;
                const int b = (Y + 2018 * U) / 1000, g = (Y - 813 * V - 391 * U) / 1000, r = (Y + 1596 * V) / 1000;
                uint8_t *o = bgr24 + (size_t)y * widthBytesBgr24 + 3 * x;
                o[0] = (uint8_t)(b < 0 ? 0 : b > 255 ? 255 : b); o[1] = (uint8_t)(g < 0 ? 0 : g > 255 ? 255 : g); o[2] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
            }
        return 0;
    }
    /* H264PictureBase.cpp:382-438: 24-bit BMP of the picture (rows bottom-up, padded to 4 bytes) */
    int saveToBmpFile(const char *filename) const {
        const uint32_t W = (uint32_t)PicWidthInSamplesL, H = (uint32_t)PicHeightInSamplesL, pitch = (W * 3 + 3) & ~3u;
        std::vector<uint8_t> bgr((size_t)pitch * H, 0), flip((size_t)pitch * H, 0);
        convertYuv420pToBgr24(W, H, m_pic_buff_luma, bgr.data(), pitch);
        for (uint32_t y = 0; y < H; y++) memcpy(&flip[(size_t)(H - 1 - y) * pitch], &bgr[(size_t)y * pitch], pitch);
        uint8_t hdr[54] = {0};
        const uint32_t size = 54 + pitch * H;
        hdr[0] = 'B'; hdr[1] = 'M'; memcpy(hdr + 2, &size, 4); hdr[10] = 54; hdr[14] = 40; memcpy(hdr + 18, &W, 4); memcpy(hdr + 22, &H, 4); hdr[26] = 1; hdr[28] = 24;
        const uint32_t img = pitch * H; memcpy(hdr + 34, &img, 4);
        FILE *f = fopen(filename, "wb");
        if (!f) return -1;
        const bool ok = fwrite(hdr, 1, 54, f) == 54 && fwrite(flip.data(), 1, flip.size(), f) == flip.size();
        fclose(f);
        return ok ? 0 : -1;
    }
};
class CH264Picture { public: CH264PictureBase m_picture_frame; };
class CH264PicturesGOP;      /* opaque here: output bumping lives in the host front end */

typedef int (*output_frame_callback)(CH264Picture *outPicture, void *userData, int errorCode);

class CH264VideoDecoder {
public:
    CH264VideoDecoder() : m_output_frame_callback(NULL), m_userData(NULL) {}
    int init() { return m_impl.init(); }
    int unInit() { return m_impl.unInit(); }
    int set_output_frame_callback_functuin(output_frame_callback cb, void *userData) {   /* sic */
        m_output_frame_callback = cb; m_userData = userData;
        return m_impl.set_output_frame_callback_functuin(&CH264VideoDecoder::trampoline, this);
    }
    int set_device(int device) { return m_impl.set_device(device); }
    int open(const char *url) { return m_impl.open(url); }
    /* H264VideoDecoder.h:42 / H264VideoDecoder.cpp:380-432: hand one picture to the registered callback (the GOP argument selected the
     * output picture in the reference; here pictures already arrive in output order, so it is unused) */
    int do_callback(CH264Picture *picture_current, CH264PicturesGOP *, int32_t) {
        if (m_output_frame_callback && picture_current) return m_output_frame_callback(picture_current, m_userData, H264_DECODE_ERROR_CODE_NO) ? -1 : 0;
        return 0;
    }
    const char *last_error() const { return m_impl.last_error(); }
private:
    static int trampoline(CH264PictureB200 *p, void *self, int errorCode) {
        CH264VideoDecoder *d = (CH264VideoDecoder *)self;
        if (!d->m_output_frame_callback) return 0;
        if (!p) return d->m_output_frame_callback(NULL, d->m_userData, errorCode);
        CH264Picture pic; memset((void *)&pic, 0, sizeof pic);
        CH264PictureBase &f = pic.m_picture_frame; const CH264PictureBaseB200 &b = p->m_picture_frame;
        f.m_pic_buff_luma = b.m_pic_buff_luma; f.m_pic_buff_cb = b.m_pic_buff_cb; f.m_pic_buff_cr = b.m_pic_buff_cr;
        f.PicWidthInSamplesL = b.PicWidthInSamplesL; f.PicHeightInSamplesL = b.PicHeightInSamplesL; f.PicWidthInSamplesC = b.PicWidthInSamplesC; f.PicHeightInSamplesC = b.PicHeightInSamplesC;
        f.PicOrderCnt = b.PicOrderCnt; f.m_PicNumCnt = b.m_PicNumCnt;
        f.m_h264_slice_header.slice_type = b.slice_type; f.m_h264_slice_header.MbaffFrameFlag = b.MbaffFrameFlag;
        f.m_h264_slice_header.m_sps.profile_idc = b.profile_idc; f.m_h264_slice_header.m_sps.level_idc = b.level_idc; f.m_h264_slice_header.m_sps.fps = b.fps;
        f.m_h264_slice_header.m_pps.entropy_coding_mode_flag = b.entropy_coding_mode_flag;
        return d->do_callback(&pic, NULL, 0);
    }
    CH264VideoDecoderB200 m_impl;
    output_frame_callback m_output_frame_callback;
    void *m_userData;
};
#endif
