// A client written the way the reference's own main.cpp is (callback that names and saves every output frame, final NULL call ends
// the run), built against include/H264VideoDecoder.h — the same-name shim — instead of the reference's headers.  tests/test_shim.py
// compiles it on the CPU box and runs it on the GPU box.  usage: shim_main_check <in.h264> <outDir>
#include "H264VideoDecoder.h"
#include <stdio.h>

static int g_frames = 0, g_end_seen = 0;

static int frame_done(CH264Picture *outPicture, void *userData, int errorCode) {
    if (!outPicture) { g_end_seen = errorCode == H264_DECODE_ERROR_CODE_FILE_END; return -1; }
    const char *outDir = (const char *)userData;
    char name[600];
    snprintf(name, sizeof name, "%s/out_%dx%d.%d.bmp", outDir, outPicture->m_picture_frame.PicWidthInSamplesL, outPicture->m_picture_frame.PicHeightInSamplesL, g_frames);
    printf("frame %d: m_PicNumCnt=%d(%s) PicOrderCnt=%d profile=%d level=%d cabac=%d fps=%.3f -> %s\n", g_frames, outPicture->m_picture_frame.m_PicNumCnt,
           H264_SLIECE_TYPE_TO_STR(outPicture->m_picture_frame.m_h264_slice_header.slice_type), outPicture->m_picture_frame.PicOrderCnt,
           outPicture->m_picture_frame.m_h264_slice_header.m_sps.profile_idc, outPicture->m_picture_frame.m_h264_slice_header.m_sps.level_idc,
           outPicture->m_picture_frame.m_h264_slice_header.m_pps.entropy_coding_mode_flag, outPicture->m_picture_frame.m_h264_slice_header.m_sps.fps, name);
    if (outPicture->m_picture_frame.saveToBmpFile(name) != 0) return -1;
    g_frames++;
    return 0;
}

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s <in.h264> <outDir>\n", argv[0]); return 2; }
    CH264VideoDecoder decoder;
    decoder.init();
    decoder.set_output_frame_callback_functuin(frame_done, argv[2]);
    const int ret = decoder.open(argv[1]);
    decoder.unInit();
    printf("RESULT ret=%d frames=%d end_seen=%d\n", ret, g_frames, g_end_seen);
    return ret != 0 || !g_end_seen;
}
