/*
 * H264VideoDecoderB200.h — host-side mirror of the reference's decoder interface
 * (jfu222/h264_video_decoder_demo, H264VideoDecoder.h:22-43) on top of the B200 reconstruction engine.
 *
 * Same shape and conventions as the reference: every method returns int, 0 = ok, non-zero = failure, no
 * exceptions; open() blocks until the end of the stream and invokes the callback synchronously on the calling
 * thread, once per output picture in display order and finally once with outPicture == NULL and
 * errorCode == H264_DECODE_ERROR_CODE_FILE_END (H264VideoDecoder.cpp:372-374).  A non-zero callback return
 * stops decoding (H264VideoDecoder.cpp:119-124).  The picture handed to the callback is owned by the decoder
 * and valid only during the callback (H264VideoDecoder.cpp:407, 431); Y, Cb and Cr are contiguous in one
 * allocation exactly like the reference's (H264PictureBase.cpp:167-179).
 *
 * open() takes an Annex-B byte stream: the serial host stage (h264_front_b200.h: NAL split, CAVLC/CABAC, every
 * neighbour-dependent derivation, DPB/output bookkeeping) runs on a producer thread and feeds per-picture
 * structure-of-arrays to the CUDA engine (h264_recon_b200.h); all pixel work happens on the GPU — there is no CPU
 * reconstruction fallback.  A pre-parsed picture container (tools/h264b2_parse, oracle/ref_harness) is accepted too.
 */
#ifndef H264_VIDEO_DECODER_B200_H
#define H264_VIDEO_DECODER_B200_H
#include <stdint.h>

enum { H264_DECODE_ERROR_CODE_NO = 0, H264_DECODE_ERROR_CODE_FILE_END = 1 };   /* H264CommonFunc.h:477-484 */

struct CH264PictureBaseB200 {          /* the fields consumers of the reference read (main.cpp:25-31) */
    uint8_t *m_pic_buff_luma, *m_pic_buff_cb, *m_pic_buff_cr;
    int32_t PicWidthInSamplesL, PicHeightInSamplesL, PicWidthInSamplesC, PicHeightInSamplesC;
    int32_t PicOrderCnt, m_PicNumCnt, slice_type, MbaffFrameFlag;
    /* m_h264_slice_header.m_sps / m_pps fields the reference's player reads (SDH264Player/MyStatic.cpp:190-211); Annex-B input only,
     * 0 for pre-parsed containers (they do not carry parameter sets) */
    int32_t profile_idc, level_idc, entropy_coding_mode_flag;
    float fps;
};
struct CH264PictureB200 { CH264PictureBaseB200 m_picture_frame; };

typedef int (*output_frame_callback_b200)(CH264PictureB200 *outPicture, void *userData, int errorCode);

class CH264VideoDecoderB200 {
public:
    CH264VideoDecoderB200();
    ~CH264VideoDecoderB200();
    int init();
    int unInit();
    int set_output_frame_callback_functuin(output_frame_callback_b200 output_frame_callback, void *userData);   /* sic: the reference's spelling */
    int set_device(int device);                 /* which GPU (default 0) */
    int open(const char *url);                  /* an Annex-B .h264 byte stream (like the reference) or a pre-parsed picture container */
    const char *last_error() const { return m_error; }
private:
    int open_bitstream(const char *url);
    output_frame_callback_b200 m_output_frame_callback;
    void *m_userData;
    int m_device;
    char m_error[512];
};

/* C entry points for non-C++ hosts (ctypes / cgo / JNI): same semantics */
extern "C" {
void *h264b2_decoder_create(void);
void h264b2_decoder_destroy(void *dec);
int h264b2_decoder_set_callback(void *dec, output_frame_callback_b200 cb, void *userData);
int h264b2_decoder_set_device(void *dec, int device);
int h264b2_decoder_open(void *dec, const char *url);
const char *h264b2_decoder_last_error(void *dec);
}
#endif
