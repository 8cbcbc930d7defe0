/*
 * h264_front_b200.h — the serial host stage in front of the B200 reconstruction engine.
 *
 * Replaces, on the host, what the reference does BEFORE its per-macroblock reconstruction calls
 * (jfu222/h264_video_decoder_demo): Annex-B NAL splitting (FileReader.cpp:127, H264NalUnit.cpp:122), SPS/PPS/slice
 * header parsing (H264SPS.cpp:222, H264PPS.cpp:144, H264SliceHeader.cpp:289), POC / reference list construction /
 * marking (H264RefPicList.cpp), the slice-data loop (H264SliceData.cpp:64-534), macroblock syntax with CAVLC and
 * CABAC (H264MacroBlock.cpp:912-1856, H264ResidualBlockCavlc.cpp, H264Cabac.cpp), and every neighbour-dependent
 * DERIVATION the reference performs inside its reconstruction functions (intra pred modes H264PictureBase.cpp:773/913,
 * motion vectors / reference indices incl. spatial direct H264InterPrediction.cpp:671-2047, prediction weights :2833,
 * reference picture selection :2117, the output bumping of H264PicturesGOP.cpp:89-165).  It emits, per picture in
 * decoding order, the structure-of-arrays of h264_recon_b200.h plus the output (display) order.  No pixel is touched here.
 *
 * Plain C ABI; every function returns int (0 ok, <0 failure) like the reference.
 */
#ifndef H264_FRONT_B200_H
#define H264_FRONT_B200_H
#include "h264_recon_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct H264B2Front H264B2Front;

/* Per-picture header; identical to the record header of the pre-parsed picture containers (oracle/ref_harness.cpp). */
typedef struct H264B2FrontPicHdr {
    int32_t decode_idx, dst_surface, clear_surface, has_inter, deblock_enable, deblock_stop_mb;
    int32_t mbaff, cqp0, cqp1, n_weights, custom_scaling, slice_type, poc, n_na;
    uint32_t n_coefs, nal_ref_idc;
    uint64_t sum_pre, sum_post;          /* 0: unknown to a parser (filled by the reference harness only) */
} H264B2FrontPicHdr;

enum { H264B2_EV_PICTURE = 1,            /* a picture is completely parsed: reconstruct it (params valid until the next call) */
       H264B2_EV_OUTPUT  = 2,            /* the reference would hand picture `decode_idx` (DPB surface `surface`) to the output callback now */
       H264B2_EV_END     = 3 };          /* end of stream */

typedef struct H264B2FrontEvent {
    int32_t kind;
    int32_t decode_idx;
    int32_t surface;
    int32_t width_mbs, height_mbs;
    H264B2FrontPicHdr hdr;               /* EV_PICTURE */
    H264B2PicParams params;              /* EV_PICTURE: host arrays inside `block` */
    void *block;                         /* EV_PICTURE: one contiguous allocation holding every array of `params` (mb_info first);
                                            owned by the caller until h264b2_front_release() */
    size_t block_bytes;
} H264B2FrontEvent;

/* alloc/free: optional allocator for the SoA buffers (e.g. h264b2_host_alloc for page-locked memory); NULL = malloc. */
typedef void *(*h264b2_front_alloc_fn)(void *user, size_t bytes);
typedef void (*h264b2_front_free_fn)(void *user, void *p);

int h264b2_front_create(H264B2Front **f, h264b2_front_alloc_fn alloc, h264b2_front_free_fn free_fn, void *user);
int h264b2_front_destroy(H264B2Front *f);
/* flags = H264B2_PACKED_COEFS | H264B2_PACKED_MOTION: pictures leave with packed coefficients / motion records (params.packed,
 * see h264b2_pack_coefs / h264b2_pack_motion): the form h264b2_submit wants for PCIe; default 0 = plain arrays (what the
 * reference parser's containers hold). */
int h264b2_front_set_packed(H264B2Front *f, int flags);
/* Annex-B byte stream, whole file (the reference also reads whole NAL units from a file buffer). The memory must stay valid. */
int h264b2_front_open_memory(H264B2Front *f, const uint8_t *data, size_t bytes);
int h264b2_front_open_file(H264B2Front *f, const char *path);
/* Closed-GOP sharding: byte offsets at which a closed GOP starts, i.e. the first parameter-set/SEI NAL in
 * front of every IDR picture (offset 0 for the first).  Returns the number of GOPs found (<= max_gops written). */
int h264b2_front_gop_offsets(const uint8_t *data, size_t bytes, size_t *offsets, int max_gops);
/* Decode only [begin, end) of the stream — one or more whole closed GOPs.  Parameter sets that precede `begin` are
 * parsed first.  more_follows = 1: the last picture of the range is completed the way the reference completes a picture that
 * is followed by another one (deblocked, H264PictureBase.cpp:707) instead of as the stream's tail picture (Q1). */
int h264b2_front_open_range(H264B2Front *f, const uint8_t *data, size_t bytes, size_t begin, size_t end, int more_follows);
/* Stream-level facts the reference keeps in the picture's slice header copy (CH264PictureBase::m_h264_slice_header.m_sps / m_pps) and its
 * second consumer reads (SDH264Player/MyStatic.cpp:190-211): valid after the first H264B2_EV_PICTURE.  fps as H264SPS.cpp:346-356
 * derives it (25 without VUI timing, else time_scale / num_units_in_tick / 2). */
typedef struct H264B2StreamInfo {
    int32_t profile_idc, level_idc, entropy_coding_mode_flag, frame_mbs_only_flag, mb_adaptive_frame_field_flag;
    int32_t width_mbs, height_mbs, max_num_ref_frames, transform_8x8_mode_flag;
    double fps;
} H264B2StreamInfo;
int h264b2_front_stream_info(H264B2Front *f, H264B2StreamInfo *info);
/* Pull the next event. Returns 0, or <0 on a fatal stream error (message in h264b2_front_last_error). */
int h264b2_front_next(H264B2Front *f, H264B2FrontEvent *ev);
/* Give a picture block back for reuse (blocks still out at destroy time are freed there). */
/* Thread safety: h264b2_front_release() may be called from another thread while h264b2_front_next() runs on the same front end
 * (the block pool is locked); every other call on one front end must come from one thread at a time. */
int h264b2_front_release(H264B2Front *f, void *block);
const char *h264b2_front_last_error(H264B2Front *f);
/* Convenience for tools/tests: parse a whole stream into a picture container file (same format the reference harness writes). */
int h264b2_front_write_container(const char *h264_path, const char *container_path, int max_pictures);
/* Same for the byte range [begin, end) of the file (a closed-GOP shard, see h264b2_front_open_range); end == 0: whole file. */
int h264b2_front_write_container_range(const char *h264_path, const char *container_path, int max_pictures, size_t begin, size_t end, int more_follows);

#ifdef __cplusplus
}
#endif
#endif
