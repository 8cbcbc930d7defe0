// h264_decoder.h — state of the host front end: DPB bookkeeping (the reference's CH264Picture / CH264PictureBase /
// CH264PicturesGOP restated as plain structs), per-macroblock transient state, the picture under construction.
#pragma once
#include "h264_front_internal.h"
#include <string>
#include <mutex>

namespace h264b2 {

const uint16_t (*coeff_token_table())[17][4][2];
const uint16_t (*total_zeros_table())[16][16][2];
const uint16_t (*run_before_table())[15][2];
const uint8_t (*me_cbp_table())[2];
const uint8_t *scan4x4_table(int field);
const uint8_t *scan8x8_table(int field);

enum MbType : uint8_t { T_NA = 0, T_I_NxN, T_I16, T_IPCM, T_SI, T_P16x16, T_P16x8, T_P8x16, T_P8x8, T_P8x8ref0, T_PSKIP,
                        T_BDIRECT, T_B16x16, T_B16x8, T_B8x16, T_B8x8, T_BSKIP };
enum { PM_NA = 0, PM_L0 = 1, PM_L1 = 2, PM_BI = 3, PM_DIRECT = 4 };       // bit0: uses list 0, bit1: uses list 1 (syntax level)
enum { MARK_UNKNOWN = 0, MARK_SHORT = 2, MARK_LONG = 3, MARK_UNUSED = 5 };   // H264_PICTURE_MARKED_AS_* (H264CommonFunc.h)

// Transient per-macroblock state of the picture being parsed; zeroed when a picture starts, exactly like the
// reference's memset of m_mbs (H264PictureBase.cpp:48-51) — several derivations depend on "never written = 0".
struct MbT {
    uint8_t type, cls, field, t8x8, skip_flag, intra, pm0_inter, ipcm;
    uint8_t cbp_luma, cbp_chroma, chroma_pred, i16mode, num_part, part_w, part_h, decoded;
    int8_t qp; int8_t qp_delta;
    uint16_t slice;
    uint8_t nnz[16];          // mb_luma_4x4_non_zero_count_coeff (incl. the I16x16 DC count landing in [0], MB:1759)
    uint8_t nnz8[4];          // mb_luma_8x8_non_zero_count_coeff
    uint8_t nnz_c[2][4];      // mb_chroma_4x4_non_zero_count_coeff (chroma DC count lands in [c][0])
    uint16_t cbf_ac[3];       // CABAC coded_block_flag per block: luma / Cb / Cr (H264Cabac.cpp:5146)
    uint8_t cbf_dc;           // bit0 luma DC, bit1 Cb DC, bit2 Cr DC
    int8_t ipred[16];         // Intra4x4PredMode[16] or Intra8x8PredMode[4]
    uint8_t part_pm[4];       // syntax-level prediction mode of the partition covering each 8x8 quadrant (PM_*)
    uint8_t sub_shape[4];     // P_8x8/B_8x8: 0 8x8, 1 8x4, 2 4x8, 3 4x4 (direct sub-macroblocks: 3)
    uint8_t sub_direct[4];
    int8_t ref_syn[2][4];     // ref_idx_lX syntax elements per quadrant
    uint8_t pf[2][4];         // m_PredFlagLX
    int8_t ref[2][4];         // m_RefIdxLX
    int16_t mvd[2][16][2];    // mvd_lX per 4x4 block (raster)
};

// What later pictures read from a reference picture's macroblocks (co-located derivation, H264InterPrediction.cpp:1243-1296).
struct ColMb { uint8_t type, intra, field, pf0; int8_t ref[2][4]; };

struct Slot {                 // one CH264Picture of the reference's 16-entry DPB (H264PicturesGOP.h:27)
    // parent level (CH264Picture)
    int p_mark = 0, p_coded_type = 0, p_coded_marked = 0, p_PicNum = 0, p_LongTermPicNum = 0, in_use = 0, p_finished = 0;
    int prev_ref = -1, prev = -1;     // m_picture_previous_ref / m_picture_previous (slot indices: the reference keeps live pointers)
    // frame level (m_picture_frame)
    int coded_type = 0, mark = 0, FrameNum = 0, FrameNumWrap = 0, PicNum = 0, LongTermPicNum = 0, LongTermFrameIdx = 0, MaxLongTermFrameIdx = -1;
    int TopFieldOrderCnt = 0, BottomFieldOrderCnt = 0, PicOrderCnt = 0, PicOrderCntMsb = 0, FrameNumOffset = 0, mmco5 = 0, mmco6 = 0;
    int hdr_poc_lsb = 0, hdr_frame_num = 0, hdr_field_pic = 0, hdr_mbaff_sps = 0;
    int slice_cnt = 0, slice_number = -1, mb_cnt = 0;
    int list[2][34]; int listlen[2] = {0, 0};
    int decode_idx = -1;
    int used_before = 0;
    std::vector<H264B2MbMotion> motion;
    std::vector<ColMb> col;
    Slot() { for (int l = 0; l < 2; l++) for (int i = 0; i < 34; i++) list[l][i] = -1; }
};

struct Nb { int mb, xW, yW; };

struct Block { uint8_t *p; size_t cap; };

struct Front {
    // configuration
    h264b2_front_alloc_fn alloc = nullptr; h264b2_front_free_fn free_fn = nullptr; void *alloc_user = nullptr;
    bool packed_coefs = false, packed_motion = false;      // h264b2_front_set_packed
    std::string error;
    // stream
    std::vector<uint8_t> file; const uint8_t *data = nullptr; size_t size = 0, nal_pos = 0;
    size_t range_begin = 0; int more_follows = 0, primed = 1;      // closed-GOP shard decoding (h264b2_front_open_range)
    std::vector<uint8_t> rbsp;
    SPS spss[32]; PPS ppss[256];
    int sps_seen = 0, pps_seen = 0, max_num_reorder_frames = 0;
    // DPB
    Slot slots[16];
    int cur = 0;                 // picture_current (slot index, -1 = NULL)
    int cur_has_ptr = 0;         // m_current_picture_ptr != NULL
    int out_buf[16]; int out_len = 0;
    int pic_num_cnt = 0, decode_count = 0;
    SliceHeader last_sh, sh;
    int eof_done = 0, stop = 0;
    // picture under construction
    int wmb = 0, hmb = 0, nmb = 0;
    std::vector<MbT> mbs;
    std::vector<H264B2MbInfo> info; std::vector<uint64_t> modes; std::vector<uint32_t> coff;
    std::vector<int16_t> coefs; size_t ncoef = 0;      // coefs is raw storage (grown ahead of the writer, never shrunk), ncoef the levels written for this picture
    std::vector<H264B2Weight> weights;
    int has_inter = 0, pic_active = 0;
    SliceHeader pic_sh;          // header of the last slice of the picture (what the reference keeps in the picture)
    // slice state
    BitReader br; Cabac cabac;
    int slice_number = 0, mb_field = 0, mb_skip_flag = 0, qp_prev = 0, CurrMbAddr = 0;
    // event queue
    std::vector<H264B2FrontEvent> events; size_t ev_pos = 0;
    std::vector<Block> free_blocks, live_blocks;
    std::mutex block_mu;         // h264b2_front_release() may run on another thread than h264b2_front_next() (decoder facade, multi-stream pipeline)
    uint8_t *get_block(size_t bytes, size_t *cap);
    void release_block(void *p);
    ~Front();

    // ---- h264_front.cpp
    int next_event(H264B2FrontEvent *ev);
    int pump();
    int handle_slice_nal(int nal_ref_idc, int nal_unit_type);
    int finish_picture(int deblock_enable);
    int end_decode_and_new_picture();
    void do_callback(int pic, int flush);
    int get_one_out(int new_pic, int *out);
    void start_picture_storage();
    int decode_poc();
    int build_ref_lists();
    void picture_numbers();
    int init_lists();
    int modify_lists();
    int mark_reference();
    void emit_picture(int deblock_enable);
    // ---- h264_slice.cpp
    int decode_slice();
};

}  // namespace h264b2
