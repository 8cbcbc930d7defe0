// h264b2_decode — the reference's main.cpp (main.cpp:67-92) against the B200 engine: decode a stream through
// CH264VideoDecoderB200::open() and write the raw I420 frames (full coded size, like the reference) in output order.
//   h264b2_decode in.bin [out.yuv]
#include "H264VideoDecoderB200.h"
#include <stdio.h>
struct Ctx { FILE *fo; int n; };
static int cb(CH264PictureB200 *pic, void *user, int errorCode) {
    Ctx *c = (Ctx *)user;
    if (!pic) { fprintf(stderr, "end of stream (errorCode %d), %d frames\n", errorCode, c->n); return 0; }
    const CH264PictureBaseB200 &f = pic->m_picture_frame;
    if (c->fo) fwrite(f.m_pic_buff_luma, 1, (size_t)f.PicWidthInSamplesL * f.PicHeightInSamplesL * 3 / 2, c->fo);   // Y|Cb|Cr contiguous
    c->n++;
    return 0;
}
int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s in.bin [out.yuv]\n", argv[0]); return 2; }
    Ctx c = { argc > 2 ? fopen(argv[2], "wb") : nullptr, 0 };
    CH264VideoDecoderB200 vd;
    vd.set_output_frame_callback_functuin(cb, &c);
    const int r = vd.open(argv[1]);
    if (r) fprintf(stderr, "open failed (%d): %s\n", r, vd.last_error());
    if (c.fo) fclose(c.fo);
    return r ? 1 : 0;
}
