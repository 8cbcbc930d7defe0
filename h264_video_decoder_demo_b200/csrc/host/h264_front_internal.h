// h264_front_internal.h — shared declarations of the host entropy/derivation stage (see include/h264_front_b200.h).
// Bit reader, parameter sets, slice header, CABAC engine.  Plain C++17, no CUDA, no oracle.
#pragma once
#include "h264_front_b200.h"
#include <stdint.h>
#include <string.h>
#include <vector>

namespace h264b2 {

// ------------------------------------------------------------------ RBSP bit reader
// Mirrors the observable behaviour of the reference's CBitstream (Bitstream.cpp:45-124): MSB first, the very last bit of a
// buffer always reads as 0 (Bitstream.cpp:52-56 — applied once when the RBSP is extracted), more_rbsp_data() as
// H264CommonFunc.cpp:71-123.  Reads past the end return 0 (the reference has undefined behaviour there).
struct BitReader {
    const uint8_t *d = nullptr;   // padded with >= 8 zero bytes
    int64_t nbits = 0, pos = 0, last1 = -1;
    void init(const uint8_t *data, size_t bytes) {
        d = data; nbits = (int64_t)bytes * 8; pos = 0; last1 = -1;
        for (int64_t i = (int64_t)bytes - 1; i >= 0; i--) if (data[i]) { int b = 0; while (!((data[i] >> b) & 1)) b++; last1 = i * 8 + (7 - b); break; }
    }
    // Past the end the reference shifts by a negative count (Bitstream.cpp:54, undefined behaviour); as compiled for x86-64
    // that re-reads the last byte once every 32 reads (24 zeros, then the 8 bits of the last byte).  Reproduced so that
    // truncated slices fail at the same syntax element (gop121's corrupt IDR).
    uint32_t peek_slow(int k) const {
        uint32_t v = 0;
        const uint32_t last = nbits ? d[(nbits >> 3) - 1] : 0;
        for (int i = 0; i < k; i++) {
            const int64_t q = pos + i;
            uint32_t b;
            if (q < nbits) b = (d[q >> 3] >> (7 - (q & 7))) & 1;
            else { const int r = (int)((q - nbits) & 31); b = r >= 24 ? (last >> (31 - r)) & 1 : 0; }
            v = (v << 1) | b;
        }
        return v;
    }
    inline uint32_t peek(int k) const {          // k in [1, 32]
        if (pos + k > nbits) return peek_slow(k);
        uint64_t v; memcpy(&v, d + (pos >> 3), 8);          // the buffer is padded, see Front::pump
        v = __builtin_bswap64(v);
        return (uint32_t)((v << (pos & 7)) >> (64 - k));
    }
    inline void skip(int k) { pos += k; }
    inline uint32_t u(int k) { if (k == 0) return 0; uint32_t v = peek(k); pos += k; return v; }
    inline uint32_t u1() { return u(1); }
    inline uint32_t ue() {
        uint32_t v = peek(32);
        if (v == 0) { pos += 32; return 0xFFFFFFFFu; }      // > 31 leading zeros: not a valid code (H264Golomb.cpp:57 fails too)
        int lz = __builtin_clz(v);
        pos += lz + 1;
        return lz ? ((1u << lz) - 1 + u(lz)) : 0;
    }
    inline int32_t se() { uint32_t k = ue(); int32_t v = (int32_t)((k + 1) >> 1); return (k & 1) ? v : -v; }
    inline uint32_t te(int range) { if (range <= 0) return 0; if (range == 1) return !u1(); return ue(); }
    inline bool more_rbsp_data() const { return pos < last1; }
    inline bool aligned() const { return (pos & 7) == 0; }
    inline bool exhausted() const { return pos >= nbits; }
};

// ------------------------------------------------------------------ parameter sets (fields the path needs)
struct SPS {
    int valid = 0;
    int profile_idc = 0, constraint_set3_flag = 0, level_idc = 0, sps_id = 0, chroma_format_idc = 1, separate_colour_plane_flag = 0;
    int bit_depth_luma_minus8 = 0, bit_depth_chroma_minus8 = 0, qpprime_y_zero_transform_bypass_flag = 0;
    int seq_scaling_matrix_present_flag = 0, seq_scaling_list_present_flag[12] = {0};
    int32_t ScalingList4x4[6][16], ScalingList8x8[6][64];
    int UseDefault4x4[6] = {0}, UseDefault8x8[6] = {0};
    int log2_max_frame_num_minus4 = 0, pic_order_cnt_type = 0, log2_max_pic_order_cnt_lsb_minus4 = 0;
    int delta_pic_order_always_zero_flag = 0, offset_for_non_ref_pic = 0, offset_for_top_to_bottom_field = 0, num_ref_frames_in_pic_order_cnt_cycle = 0;
    int offset_for_ref_frame[256];
    int max_num_ref_frames = 0, gaps_in_frame_num_value_allowed_flag = 0, pic_width_in_mbs_minus1 = 0, pic_height_in_map_units_minus1 = 0;
    int frame_mbs_only_flag = 1, mb_adaptive_frame_field_flag = 0, direct_8x8_inference_flag = 0;
    int max_num_reorder_frames = -1;
    int timing_info_present_flag = 0; uint32_t num_units_in_tick = 0, time_scale = 0;      // VUI timing (the reference derives its `fps` from it)
    // derived
    int PicWidthInMbs = 0, PicHeightInMapUnits = 0, FrameHeightInMbs = 0, ChromaArrayType = 1, MaxFrameNum = 16, MaxPicOrderCntLsb = 16, ExpectedDeltaPerPicOrderCntCycle = 0;
    SPS() { memset(ScalingList4x4, 0, sizeof ScalingList4x4); memset(ScalingList8x8, 0, sizeof ScalingList8x8); memset(offset_for_ref_frame, 0, sizeof offset_for_ref_frame); }
};

struct PPS {
    int valid = 0;
    int pps_id = 0, sps_id = 0, entropy_coding_mode_flag = 0, bottom_field_pic_order_in_frame_present_flag = 0, num_slice_groups_minus1 = 0;
    int num_ref_idx_l0_default_active_minus1 = 0, num_ref_idx_l1_default_active_minus1 = 0, weighted_pred_flag = 0, weighted_bipred_idc = 0;
    int pic_init_qp_minus26 = 0, pic_init_qs_minus26 = 0, chroma_qp_index_offset = 0, deblocking_filter_control_present_flag = 0;
    int constrained_intra_pred_flag = 0, redundant_pic_cnt_present_flag = 0, transform_8x8_mode_flag = 0, pic_scaling_matrix_present_flag = 0;
    int pic_scaling_list_present_flag[12] = {0};
    int32_t ScalingList4x4[6][16], ScalingList8x8[6][64];
    int UseDefault4x4[6] = {0}, UseDefault8x8[6] = {0};
    int second_chroma_qp_index_offset = 0;
    PPS() { memset(ScalingList4x4, 0, sizeof ScalingList4x4); memset(ScalingList8x8, 0, sizeof ScalingList8x8); }
};

enum { SLICE_P = 0, SLICE_B = 1, SLICE_I = 2, SLICE_SP = 3, SLICE_SI = 4 };

struct Mmco { int op = 0, difference_of_pic_nums_minus1 = 0, long_term_pic_num = 0, long_term_frame_idx = 0, max_long_term_frame_idx_plus1 = 0; };

struct SliceHeader {
    int nal_ref_idc = 0, nal_unit_type = 0, IdrPicFlag = 0;
    int first_mb_in_slice = 0, slice_type = 0, pps_id = 0, colour_plane_id = 0, frame_num = 0, field_pic_flag = 0, bottom_field_flag = 0, idr_pic_id = 0;
    int pic_order_cnt_lsb = 0, delta_pic_order_cnt_bottom = 0, delta_pic_order_cnt[2] = {0, 0}, redundant_pic_cnt = 0, direct_spatial_mv_pred_flag = 0;
    int num_ref_idx_active_override_flag = 0, num_ref_idx_l0_active_minus1 = 0, num_ref_idx_l1_active_minus1 = 0;
    int ref_pic_list_modification_flag[2] = {0, 0}, modification_count[2] = {0, 0};
    int modification_of_pic_nums_idc[2][33], abs_diff_pic_num_minus1[2][33], long_term_pic_num[2][33];
    int luma_log2_weight_denom = 0, chroma_log2_weight_denom = 0;
    int last_luma_weight_flag[2] = {0, 0};      // luma_weight_lX_flag of the last entry parsed: what luma_weight_lX[-1] reads in the reference (Q8)
    int luma_weight[2][32], luma_offset[2][32], chroma_weight[2][32][2], chroma_offset[2][32][2];
    int no_output_of_prior_pics_flag = 0, long_term_reference_flag = 0, adaptive_ref_pic_marking_mode_flag = 0, mmco_count = 0;
    Mmco mmco[33];
    int cabac_init_idc = 0, slice_qp_delta = 0, disable_deblocking_filter_idc = 0, slice_alpha_c0_offset_div2 = 0, slice_beta_offset_div2 = 0;
    // derived
    int SliceQPY = 0, MbaffFrameFlag = 0, PicHeightInMbs = 0, PicSizeInMbs = 0, MaxPicNum = 0, CurrPicNum = 0, FilterOffsetA = 0, FilterOffsetB = 0;
    int32_t ScalingList4x4[6][16], ScalingList8x8[6][64];
    SPS sps; PPS pps;               // snapshots, like the reference's m_sps / m_pps copies (H264SliceHeader.cpp:300-301)
    SliceHeader() {
        memset(modification_of_pic_nums_idc, 0, sizeof modification_of_pic_nums_idc); memset(abs_diff_pic_num_minus1, 0, sizeof abs_diff_pic_num_minus1);
        memset(long_term_pic_num, 0, sizeof long_term_pic_num); memset(luma_weight, 0, sizeof luma_weight); memset(luma_offset, 0, sizeof luma_offset);
        memset(chroma_weight, 0, sizeof chroma_weight); memset(chroma_offset, 0, sizeof chroma_offset);
        memset(ScalingList4x4, 0, sizeof ScalingList4x4); memset(ScalingList8x8, 0, sizeof ScalingList8x8);
    }
};

int parse_sps(BitReader &br, SPS &sps);
int parse_pps(BitReader &br, PPS &pps, const SPS *spss);
int parse_slice_header(BitReader &br, int nal_ref_idc, int nal_unit_type, const SPS *spss, const PPS *ppss, SliceHeader &sh);
bool first_vcl_nal_of_picture(const SliceHeader &cur, const SliceHeader &last);

// ------------------------------------------------------------------ CABAC arithmetic decoder (9.3.1.2, 9.3.3.2; H264Cabac.cpp:1041-1086, 2577-2824)
extern const uint8_t (&g_range_lps)[64][4];
extern uint8_t g_next_state[2][128];      // [0 = MPS decoded, 1 = LPS decoded][(pStateIdx << 1) | valMPS]  (Table 9-45 on the packed state)
struct Cabac {
    BitReader *br = nullptr;
    uint32_t range = 0, offset = 0;
    uint8_t state[1024];          // (pStateIdx << 1) | valMPS; 16 bit on purpose: a byte store may alias range / offset / win and would force them through memory after every decision
    void init_contexts(int slice_type, int cabac_init_idc, int slice_qp);
    // The decoder keeps codIOffset together with the next `k` bits of the stream: V = codIOffset * 2^k + (those k bits), so
    // "codIOffset < codIRange" is "V < codIRange << k" and a renormalisation shift is just k -= n (no bit is moved).  32 bits are
    // loaded at a time (BitReader::peek is a pure function of the position, past-the-end behaviour included, so reading ahead
    // changes nothing); br->pos runs k bits ahead of the reference's read position and is put back by sync() before anyone else
    // reads the stream (I_PCM samples, re-initialisation).  `offset` is only kept for init / diagnostics.
    uint64_t V = 0; int k = 0;
    inline void sync() { br->pos -= k; V >>= k; k = 0; }
    __attribute__((noinline)) void refill() { V = (V << 32) | br->peek(32); br->pos += 32; k += 32; }
    void init_engine(BitReader *b) { if (br && k) sync(); br = b; range = 510; offset = br->u(9); V = offset; k = 0; }      // re-initialisation inside a slice (I_PCM)
    void start_slice(BitReader *b) { br = b; range = 510; offset = br->u(9); V = offset; k = 0; }                          // the look-ahead of the previous slice is void
    __attribute__((always_inline)) inline int decision(int ctx) {
        // branch-free except for the refill (once per 32 bits): which symbol was decoded and whether a renormalisation follows are
        // both close to coin flips for the coefficient contexts, a mispredicted branch costs more than the whole arithmetic
        const uint32_t s = state[ctx];
        const uint32_t rlps = g_range_lps[s >> 1][(range >> 6) & 3];
        const uint32_t rmps = range - rlps;
        const uint64_t scaled = (uint64_t)rmps << k;
        const uint32_t lps = V >= scaled;                 // 1: least probable symbol
        V -= lps ? scaled : 0;
        uint32_t r = lps ? rlps : rmps;
        state[ctx] = (&g_next_state[0][0])[(lps << 7) | s];      // the table as one array of 256: [1][s] follows [0][127]
        const int n = __builtin_clz(r) - 23;               // r in [6, 510] -> 0..6 (0 or 1 after an MPS)
        range = r << n;
        if (k < n) refill();
        k -= n;
        return (int)((s & 1) ^ lps);
    }
    inline int bypass() {
        if (k == 0) refill();
        k--;
        const uint64_t scaled = (uint64_t)range << k;
        const uint32_t b = V >= scaled;          // signs and suffix bits are coin flips: no branch
        V -= b ? scaled : 0;
        return (int)b;
    }
    inline int terminate() {
        range -= 2;
        if (V >= ((uint64_t)range << k)) return 1;
        if (range < 256) { const int n = __builtin_clz(range) - 23; range <<= n; if (k < n) refill(); k -= n; }
        return 0;
    }
};

// The decoder's registers as locals of one syntax function (the residual loops): `state` is a byte array, so a store into it may alias
// range / V / k of the Cabac object and the compiler keeps them in memory; a CabacRegs whose address never escapes lives in
// registers.  Same arithmetic as Cabac::decision / bypass; written back by the destructor.
struct CabacRegs {
    Cabac &c; uint32_t range; uint64_t V; int k; uint8_t *const st; const uint8_t (*const lps)[4]; const uint8_t *const next;
    explicit CabacRegs(Cabac &cc) : c(cc), range(cc.range), V(cc.V), k(cc.k), st(cc.state), lps(g_range_lps), next(&g_next_state[0][0]) {}
    ~CabacRegs() { c.range = range; c.V = V; c.k = k; }
    __attribute__((noinline)) void refill() { V = (V << 32) | c.br->peek(32); c.br->pos += 32; k += 32; }
    __attribute__((always_inline)) inline int decision(int ctx) {
        const uint32_t s = st[ctx];
        const uint32_t rlps = lps[s >> 1][(range >> 6) & 3];
        const uint32_t rmps = range - rlps;
        const uint64_t scaled = (uint64_t)rmps << k;
        const uint32_t l = V >= scaled;
        V -= l ? scaled : 0;
        const uint32_t r = l ? rlps : rmps;
        st[ctx] = next[(l << 7) | s];
        const int n = __builtin_clz(r) - 23;
        range = r << n;
        if (__builtin_expect(k < n, 0)) refill();
        k -= n;
        return (int)((s & 1) ^ l);
    }
    __attribute__((always_inline)) inline int bypass() {
        if (__builtin_expect(k == 0, 0)) refill();
        k--;
        const uint64_t scaled = (uint64_t)range << k;
        const uint32_t b = V >= scaled;
        V -= b ? scaled : 0;
        return (int)b;
    }
};

}  // namespace h264b2
