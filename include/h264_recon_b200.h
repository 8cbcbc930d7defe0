/*
 * h264_recon_b200.h — C ABI of the B200 H.264 picture-reconstruction engine.
 *
 * This is the drop-in boundary for the hot path of jfu222/h264_video_decoder_demo:
 * everything the reference does per macroblock AFTER entropy decoding and
 * motion/mode derivation, i.e.
 *
 *   - residual:  dequant + 4x4 / 8x8 / DC inverse transforms
 *                (reference: H264PictureBase.cpp:3401-3929, 3989-4402, 4993-5088,
 *                 H264InterPrediction.cpp:22-407)
 *   - inter:     luma 6-tap / chroma bilinear MC, default/explicit/implicit
 *                weighting (H264InterPrediction.cpp:412-667, 2051-2829)
 *   - intra:     4x4 / 8x8 / 16x16 / chroma prediction (H264PictureBase.cpp:1062-2499)
 *   - deblock:   bS derivation + edge filters (H264PictureDeblockingFilterProcess.cpp:76-1522)
 *   - DPB pixel storage (H264PictureBase.cpp:128-233; zero-at-reuse :45-69)
 *
 * The reference has no FFI for this path (it is reached by member calls from
 * CH264SliceData::slice_data, H264SliceData.cpp:222-227, 333-338, 400-487 and
 * from end_decode_the_picture_and_get_a_new_empty_picture, H264PictureBase.cpp:707).
 * The host side (serial entropy decode + derivations) replaces those calls by
 * "append this macroblock to the picture's structure-of-arrays" and hands the
 * finished picture to h264b2_submit().  Plain pointers and sizes only.
 *
 * Conventions follow the reference: every function returns int, 0 = ok,
 * negative = failure; no exceptions; one context per GPU; calls on one
 * context must come from one thread at a time.
 */
#ifndef H264_RECON_B200_H
#define H264_RECON_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H264B2_ABI_VERSION 2

/* ---- macroblock classes (derived from CH264MacroBlock::m_mb_pred_mode /
 *      m_name_of_mb_type, H264MacroBlock.h:210-211) ---- */
enum {
    H264B2_MB_NA    = 0,  /* never decoded (MB_TYPE_NA): pixels stay 0, deblock stops here (Q2) */
    H264B2_MB_I4x4  = 1,  /* Intra_4x4   */
    H264B2_MB_I8x8  = 2,  /* Intra_8x8   */
    H264B2_MB_I16x16= 3,  /* Intra_16x16 */
    H264B2_MB_IPCM  = 4,  /* I_PCM (pred mode Intra_NA: NOT "intra" for bS, Q11) */
    H264B2_MB_INTER = 5   /* every P_* / B_* type incl. skip and direct */
};

/* H264B2MbInfo.flags */
#define H264B2_MBF_FIELD        0x01  /* mb_field_decoding_flag (MBAFF field macroblock) */
#define H264B2_MBF_T8x8         0x02  /* transform_size_8x8_flag */
#define H264B2_MBF_SPSI         0x04  /* slice_type of the MB's slice is SP or SI (bS rules) */
#define H264B2_MBF_CIP_UNAVAIL  0x08  /* IS_INTER_Prediction_Mode(mode) && constrained_intra_pred_flag:
                                         samples of this MB are "not available" to intra neighbours
                                         (H264PictureBase.cpp:1129, 1492, 1905, 2158; Q12) */

/* coef_mask bits: which coefficient blocks of the MB are present in the
 * coefficient stream (blocks that are absent are all-zero).  Blocks are stored
 * back to back, in this bit order, starting at coef_offset (units: int16). */
#define H264B2_CM_LUMA(b)   (1u << (b))      /* b=0..15: luma 4x4 block b (luma4x4BlkIdx), 16 int16 in
                                                list order (LumaLevel4x4[b][k]; for Intra16x16 k=0 is
                                                unused and k=1..15 = Intra16x16ACLevel[b][k-1]).
                                                With T8x8: b=0..3 = LumaLevel8x8[b][0..63], 64 int16. */
#define H264B2_CM_LUMA_DC   (1u << 16)       /* Intra16x16DCLevel[0..15], 16 int16 */
#define H264B2_CM_CHROMA_DC (1u << 17)       /* ChromaDCLevel[0][0..3] then [1][0..3], 8 int16 */
#define H264B2_CM_CB(b)     (1u << (18 + (b))) /* b=0..3: 16 int16, k=0 unused, k=1..15 = ChromaACLevel[0][b][k-1] */
#define H264B2_CM_CR(b)     (1u << (22 + (b))) /* b=0..3: same for Cr */
#define H264B2_CM_PCM       (1u << 26)       /* I_PCM: 384 samples stored as 384 int16 (256 Y, 64 Cb, 64 Cr) */

typedef struct H264B2MbInfo {      /* 16 bytes, one per macroblock address */
    uint8_t  mb_class;             /* H264B2_MB_* */
    uint8_t  flags;                /* H264B2_MBF_* */
    uint8_t  pred16_chroma;        /* bits0-1 Intra16x16PredMode, bits2-3 intra_chroma_pred_mode */
    int8_t   qpy;                  /* QPY (== QP'Y at 8 bit); the value deblocking reads (DB:898) */
    uint16_t slice_number;         /* CH264MacroBlock::slice_number: neighbour availability (PB:2890) */
    uint16_t nnz_mask;             /* bit b: transform block covering luma4x4BlkIdx b has non-zero
                                      levels, exactly as DB:1124-1143 tests it (Q18) */
    int8_t   filter_offset_a;      /* FilterOffsetA of the MB's slice (DB:884) */
    int8_t   filter_offset_b;
    uint8_t  deblock_idc;          /* disable_deblocking_filter_idc of the MB's slice */
    uint8_t  reserved;
    uint32_t coef_mask;            /* H264B2_CM_* */
} H264B2MbInfo;

typedef struct H264B2MbMotion {    /* 152 bytes, one per macroblock address (inter MBs only are read) */
    int16_t mv[2][16][2];          /* [list][4x4 block in RASTER order y*4+x][x,y], quarter-pel
                                      (flattening of m_MvL0/L1[mbPart][subMbPart], IP:593-597) */
    int8_t  ref_surf[2][4];        /* [list][8x8 quadrant]: -1 = list unused (predFlag 0), else
                                      (dpb_surface << 2) | view, view 0 frame / 1 top field / 2 bottom
                                      field: the result of Reference_picture_selection_process (IP:2117) */
    int8_t  ref_ident[2][4];       /* picture identity for bS "same reference picture" tests, computed
                                      the way DB:1175-1178 does (raw refIdx into the picture's last-built
                                      list, Q6); -1 = NULL */
    uint16_t wt_idx[4];            /* [8x8 quadrant] index into the picture's weight table */
} H264B2MbMotion;

typedef struct H264B2Weight {      /* 32 bytes; entry 0 of every table must be the default entry */
    int16_t mode;                  /* 0 = default weighted prediction (IP:2617), 1 = explicit/implicit formula (IP:2699) */
    int16_t logwd[3];              /* [Y,Cb,Cr] */
    int16_t w0[3], w1[3];          /* weight applied to the list-0 / list-1 prediction */
    int16_t o0[3], o1[3];          /* offsets; for single-list prediction from list 1 the device uses
                                      w1/o1, so the host puts whatever the reference would use there (Q8) */
} H264B2Weight;

typedef struct H264B2PicParams {
    int32_t  width_mbs;            /* PicWidthInMbs */
    int32_t  height_mbs;           /* frame height in MBs (68 for 1088) */
    int32_t  mbaff_frame_flag;     /* MbaffFrameFlag; field pictures (PAFF) are rejected (Q16) */
    int32_t  chroma_qp_offset[2];  /* chroma_qp_index_offset, second_chroma_qp_index_offset */
    int32_t  dst_surface;          /* DPB surface this picture is reconstructed into */
    int32_t  clear_surface;        /* 1: zero the surface first (PB:53-69); required when some MB is NA */
    int32_t  has_inter;            /* 0: no inter MB in the picture (motion may be NULL) */
    int32_t  deblock_enable;       /* 0: leave the picture un-deblocked (last picture of a stream, Q1) */
    int32_t  deblock_stop_mb;      /* deblock MB addresses [0, stop) only (first NA MB, Q2) */
    int32_t  n_weights;            /* entries in weights[] (>= 1) */
    uint32_t n_coefs;              /* int16 elements in coefs[] */
    int32_t  custom_scaling;       /* 0: Flat_4x4_16 / Flat_8x8_16; 1: level_scale4/8 given */
    int32_t  packed;               /* h264b2_submit only, bit mask.  H264B2_PACKED_COEFS: coefs points at a blob written by
                                      h264b2_pack_coefs() for the n_coefs levels; H264B2_PACKED_MOTION: motion points at a blob
                                      written by h264b2_pack_motion() for the n_mbs records.  0: plain arrays */
    /* host (or device, see h264b2_submit_device) arrays.  Alignment: the engine keeps each array's address modulo 256 when it
     * copies it to the device and its kernels use 16-byte loads, so mb_info / motion / weights / coefs must start on 16-byte
     * boundaries (intra_modes 8, coef_offset 4); h264b2_host_alloc and the front end's picture blocks guarantee it. */
    const H264B2MbInfo   *mb_info;     /* [width_mbs*height_mbs] */
    const uint64_t       *intra_modes; /* [n_mbs] 16 x 4 bit: Intra4x4PredMode[b] (b=luma4x4BlkIdx) or
                                          Intra8x8PredMode[b] in nibbles 0..3 */
    const uint32_t       *coef_offset; /* [n_mbs] int16 index of the MB's first coefficient block */
    const H264B2MbMotion *motion;      /* [n_mbs] */
    const H264B2Weight   *weights;     /* [n_weights] */
    const int16_t        *coefs;       /* [n_coefs] */
    const int16_t        *level_scale4;/* optional [2 intra/inter][2 frame/field scan][6][16] (list order k) */
    const int16_t        *level_scale8;/* optional [2][2][6][64] */
} H264B2PicParams;

typedef struct H264B2Context H264B2Context;

/* Create a per-GPU context: n_streams independent streams, each with
 * surfaces_per_stream DPB surfaces (the reference keeps 16 pictures,
 * H264PicturesGOP.h:27) of width_mbs x height_mbs macroblocks, I420,
 * Y|Cb|Cr contiguous in one allocation like PB:167-179. */
int h264b2_create(H264B2Context **ctx, int device, int n_streams, int surfaces_per_stream,
                  int width_mbs, int height_mbs);
int h264b2_destroy(H264B2Context *ctx);

/* Reconstruct one picture per listed stream (pictures of one stream are serial;
 * pictures of different streams run concurrently in one launch sequence).
 * Host arrays are DMA'd to the device asynchronously on a copy stream (one transfer per contiguous
 * span; pictures whose arrays lie back to back need one transfer) into an H264B2_SUBMIT_DEPTH-deep device arena, so the
 * copy of batch k+1 overlaps the kernels of batch k.  The call returns after enqueueing; the host
 * arrays must stay valid until h264b2_sync() or until H264B2_SUBMIT_DEPTH further submits have RETURNED (a submit waits
 * for the transfer issued H264B2_SUBMIT_DEPTH submits earlier before it reuses that arena slot).  stream_ids[i] in [0, n_streams). */
#define H264B2_SUBMIT_DEPTH 4
int h264b2_submit(H264B2Context *ctx, int n_pics, const int32_t *stream_ids,
                  const H264B2PicParams *pics);

/* Packed coefficient transport ("sparse coefficient packing to cut PCIe traffic").  The levels the
 * reference keeps in CH264MacroBlock::LumaLevel4x4 / LumaLevel8x8 / ChromaACLevel ... (H264MacroBlock.h:205-216) are almost all
 * zero after entropy decoding (5.7 % non-zero in the bundled 1080p streams), so instead of the dense coefs[] the host may send
 * one 16-bit significance map per chunk of 16 levels plus the non-zero levels (1.96 MB -> 0.25 MB per picture); the engine
 * rebuilds the dense array in HBM (k_expand) before the residual kernel runs.  coef_offset[] / n_coefs / coef_mask keep their
 * dense meaning.  h264b2_pack_coefs writes the blob (out: 16-byte aligned, cap >= h264b2_pack_coefs_bound(n_coefs) always
 * suffices; *bytes = blob size, a multiple of 16); set packed |= H264B2_PACKED_COEFS and coefs = blob.  The pack/unpack
 * functions are plain host functions (no GPU needed). */
#define H264B2_PACKED_COEFS  1
#define H264B2_PACKED_MOTION 2
size_t h264b2_pack_coefs_bound(uint32_t n_coefs);
int h264b2_pack_coefs(const int16_t *dense, uint32_t n_coefs, void *out, size_t cap, size_t *bytes);
int h264b2_unpack_coefs(const void *packed, int16_t *dense, uint32_t n_coefs);
/* Motion records travel the same way: inside every record the 16 vectors of a list are XORed with their predecessor (motion
 * is far coarser than 4x4: 81 % of the inter macroblocks of the bundled streams carry one vector per list, so 15 of 16 words
 * become 0), then the records are packed as one int16 stream (1.24 MB -> ~0.4 MB per 1080p picture).  The engine undoes both
 * steps in HBM (k_expand, k_unmotion).  Bound: h264b2_pack_coefs_bound(n_mbs * 76). */
int h264b2_pack_motion(const H264B2MbMotion *motion, uint32_t n_mbs, void *out, size_t cap, size_t *bytes);
int h264b2_unpack_motion(const void *packed, H264B2MbMotion *motion, uint32_t n_mbs);

/* Same, but every array pointer in pics[] is already a device pointer
 * (pre-parsed buffers resident in HBM: the replay path the bench's `value` times). */
int h264b2_submit_device(H264B2Context *ctx, int n_pics, const int32_t *stream_ids,
                         const H264B2PicParams *pics);

/* Copy a reconstructed surface to host I420 (width_mbs*16 x height_mbs*16, 3/2 bytes per pixel).
 * Synchronises with all prior work on the context. */
int h264b2_read_picture(H264B2Context *ctx, int stream_id, int surface, uint8_t *host_i420);

/* Asynchronous read-back of n surfaces (one per listed stream, n <= n_streams) into host memory
 * (pinned memory from h264b2_host_alloc gives true overlap): the surfaces are snapshotted on the launch
 * stream, then drained over PCIe on a separate stream, so the next submit does not wait for the copy.
 * host[i] is complete after h264b2_sync().  This is what the output callback path uses: the reference hands
 * the callback a host picture (H264VideoDecoder.cpp:118, 396-432). */
int h264b2_read_pictures_async(H264B2Context *ctx, int n, const int32_t *stream_ids, const int32_t *surfaces,
                               uint8_t *const *host_i420);

/* Output stage: the reference's integer BT.601 YUV420P -> BGR24 conversion
 * (H264PictureBase.cpp:440-468; flip_lines = 1: the bottom-up twin :471-498 used for BMP files) done on the GPU,
 * result copied to host_bgr24 (width_bytes >= 3 * width; rows beyond 3*width bytes are zero).  Synchronous. */
int h264b2_read_picture_bgr24(H264B2Context *ctx, int stream_id, int surface, uint8_t *host_bgr24, int width_bytes, int flip_lines);

/* Write a surface from host I420 (used by tests to seed reference pictures). */
int h264b2_write_picture(H264B2Context *ctx, int stream_id, int surface, const uint8_t *host_i420);

/* 64-bit positional checksum of a surface computed on the GPU (the only data
 * that has to leave the GPU in the multi-stream/multi-GPU configuration). */
int h264b2_checksum_picture(H264B2Context *ctx, int stream_id, int surface, uint64_t *checksum);

int h264b2_checksum_pictures(H264B2Context *ctx, int n, const int32_t *stream_ids, const int32_t *surfaces, uint64_t *checksums);

/* Device pointer of a surface (Y plane; Cb = Y + W*H, Cr = Cb + W*H/4). */
int h264b2_surface_ptr(H264B2Context *ctx, int stream_id, int surface, void **dev_ptr);

/* Device-memory helpers so that non-CUDA hosts can make buffers resident. */
int h264b2_dev_alloc(H264B2Context *ctx, size_t bytes, void **dev_ptr);
int h264b2_dev_free(H264B2Context *ctx, void *dev_ptr);
int h264b2_dev_upload(H264B2Context *ctx, void *dev_dst, const void *host_src, size_t bytes);
int h264b2_dev_copy(H264B2Context *ctx, void *dev_dst, const void *dev_src, size_t bytes);
/* Page-locked host memory: SoA buffers the host entropy stage fills here are DMA'd by h264b2_submit without a staging copy. */
int h264b2_host_alloc(H264B2Context *ctx, size_t bytes, void **host_ptr);
int h264b2_host_free(H264B2Context *ctx, void *host_ptr);

/* Look-ahead (default on; H264B2_LOOKAHEAD=0 in the environment turns it off at creation): descriptor upload, unpacking, the
 * residual kernel and the boundary-strength kernel of a submit depend on nothing the previous submit produces, so they are
 * enqueued on a second stream and run while the previous batch's dependency-bound wavefront kernels leave the SMs idle.
 * on = 0 serialises every kernel on the launch stream (what per-kernel timings and profiles want). */
int h264b2_set_lookahead(H264B2Context *ctx, int on);

/* Diagnostic (host only, no GPU): the (index, index, index, two-tap) tables the intra kernel uses for the directional modes 3..8 of
 * Intra_4x4 (n = 4) / Intra_8x8 (n = 8), out[6][n*n]; layout in csrc/intra.cuh (pred_tab_px).  Lets a CPU test hold them against
 * the prediction equations (H264PictureBase.cpp:1174-1395, 1606-1831). */
int h264b2_debug_intra_tables(int n, uint16_t *out);

/* Block until all enqueued work is done. */
int h264b2_sync(H264B2Context *ctx);

/* Event timing on the context's launch stream (CUDA events; for bench.py). */
int h264b2_timer_start(H264B2Context *ctx);
int h264b2_timer_stop(H264B2Context *ctx, float *elapsed_ms);
/* Per-kernel-class accumulated device time (ms) and launch counts of the last timed region, 6 entries each:
 * [0] clears/memsets, [1] k_inter (MC + residual add), [2] k_intra, [3] k_bs, [4] k_deblock, [5] k_residual. */
int h264b2_kernel_times(H264B2Context *ctx, float *ms6, int64_t *launches6);

/* ABI version / last error string (static storage). */
int h264b2_abi_version(void);
const char *h264b2_last_error(void);

/* checksum used throughout: sum over little-endian u32 words w_i of (w_i + 1) * ((2*i+1) * 0x9E3779B97F4A7C15) mod 2^64 */
uint64_t h264b2_checksum_host(const uint8_t *data, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* H264_RECON_B200_H */
