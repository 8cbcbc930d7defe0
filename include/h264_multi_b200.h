/*
 * h264_multi_b200.h — many Annex-B streams through one GPU: the serial host stage (h264_front_b200.h) of each stream runs
 * on a pool of parser threads, the reconstruction of one picture of every ready stream goes to the CUDA engine as ONE batched
 * h264b2_submit() (h264_recon_b200.h), output pictures come back in each stream's display order exactly when the reference
 * decoder would hand them to its callback (H264VideoDecoder.cpp:380-432).  This is the multi-stream form of
 * CH264VideoDecoder::open(): streams are independent, so there is no exchange between them.
 */
#ifndef H264_MULTI_B200_H
#define H264_MULTI_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct H264B2MultiStats {
    double seconds;             /* wall clock: first parser thread started -> last output picture complete (device synchronised) */
    double parse_seconds;       /* summed busy time of the parser threads (host entropy/derivation stage) */
    int64_t pictures;           /* pictures reconstructed */
    int64_t frames_out;         /* pictures delivered in output order */
    int64_t h2d_bytes;          /* structure-of-arrays bytes handed to h264b2_submit */
    int64_t d2h_bytes;          /* picture bytes copied back to page-locked host memory */
    int32_t threads, streams, submits, width_mbs, height_mbs;
    int32_t units;              /* decoding units = streams, or closed-GOP shards with H264B2_MULTI_SPLIT_GOPS */
    int32_t reserved;
} H264B2MultiStats;

enum { H264B2_MULTI_READBACK = 1,     /* copy every output picture to page-locked host memory (what an output callback receives) */
       H264B2_MULTI_SPLIT_GOPS = 2 }; /* decode the closed GOPs of every stream as independent units (own DPB each), SURVEY 8(e) */

/* flags: H264B2_MULTI_*; without READBACK pictures stay on the GPU (only 64-bit checksums leave it when stream_hash != NULL).
 * stream_hash[n_streams] (optional): per stream, h = h * 0x100000001B3 + checksum(frame) over its output frames in output order
 *           (checksums computed on the GPU; the same chain oracle/ref_harness prints as stream_hash).
 * Returns 0, or <0 with a message in err. */
int h264b2_multi_decode(int device, int n_streams, const char *const *paths, int n_threads, int flags,
                        uint64_t *stream_hash, H264B2MultiStats *stats, char *err, size_t err_len);

#ifdef __cplusplus
}
#endif
#endif
