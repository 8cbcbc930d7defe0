// h264_params.cpp — SPS / PPS / slice header syntax (7.3.2.1, 7.3.2.2, 7.3.3) and CABAC context initialisation (9.3.1.1).
// Behaviour follows the reference: H264SPS.cpp:222-457, H264VUI.cpp:104-179, H264PPS.cpp:144-235, H264SliceHeader.cpp:289-722,
// 928-1160 (scaling-list fall-back incl. its deviations), 1161-1207; H264Cabac.cpp:1041-1075.
#include "h264_front_internal.h"
#include "h264_tables.inc"

namespace h264b2 {

static void scaling_list(BitReader &br, int32_t *list, int n, int &use_default) {      // H264CommonFunc.cpp:48-68
    int last = 8, next = 8;
    for (int j = 0; j < n; j++) {
        if (next != 0) { int delta = br.se(); next = (last + delta + 256) % 256; use_default = (j == 0 && next == 0); }
        list[j] = (next == 0) ? last : next;
        last = list[j];
    }
}

static void hrd_parameters(BitReader &br) {
    int cpb_cnt_minus1 = (int)br.ue(); br.u(4); br.u(4);
    for (int i = 0; i <= cpb_cnt_minus1 && i < 32; i++) { br.ue(); br.ue(); br.u1(); }
    br.u(5); br.u(5); br.u(5); br.u(5);
}

static const int kLevelMaxDpbMbs[19][2] = {   // Table A-1 (MaxDpbMbs), the levels the reference knows (H264SPS.cpp:20-41)
    {10, 396}, {11, 900}, {12, 2376}, {13, 2376}, {20, 2376}, {21, 4752}, {22, 8100}, {30, 8100}, {31, 18000}, {32, 20480},
    {40, 32768}, {41, 32768}, {42, 34816}, {50, 110400}, {51, 184320}, {52, 184320}, {60, 696320}, {61, 696320}, {62, 696320}};

int parse_sps(BitReader &br, SPS &s) {
    s = SPS();
    s.profile_idc = br.u(8);
    br.u(3); s.constraint_set3_flag = br.u1(); br.u(2); br.u(2);
    s.level_idc = br.u(8);
    s.sps_id = (int)br.ue();
    const int p = s.profile_idc;
    if (p == 100 || p == 110 || p == 122 || p == 244 || p == 44 || p == 83 || p == 86 || p == 118 || p == 128 || p == 138 || p == 139 || p == 134 || p == 135) {
        s.chroma_format_idc = (int)br.ue();
        if (s.chroma_format_idc == 3) s.separate_colour_plane_flag = br.u1();
        s.bit_depth_luma_minus8 = (int)br.ue(); s.bit_depth_chroma_minus8 = (int)br.ue();
        s.qpprime_y_zero_transform_bypass_flag = br.u1();
        s.seq_scaling_matrix_present_flag = br.u1();
        if (s.seq_scaling_matrix_present_flag)
            for (int i = 0; i < (s.chroma_format_idc != 3 ? 8 : 12); i++) {
                s.seq_scaling_list_present_flag[i] = br.u1();
                if (s.seq_scaling_list_present_flag[i]) { if (i < 6) scaling_list(br, s.ScalingList4x4[i], 16, s.UseDefault4x4[i]); else scaling_list(br, s.ScalingList8x8[i - 6], 64, s.UseDefault8x8[i - 6]); }
            }
    }
    s.log2_max_frame_num_minus4 = (int)br.ue();
    s.pic_order_cnt_type = (int)br.ue();
    if (s.pic_order_cnt_type == 0) s.log2_max_pic_order_cnt_lsb_minus4 = (int)br.ue();
    else if (s.pic_order_cnt_type == 1) {
        s.delta_pic_order_always_zero_flag = br.u1(); s.offset_for_non_ref_pic = br.se(); s.offset_for_top_to_bottom_field = br.se();
        s.num_ref_frames_in_pic_order_cnt_cycle = (int)br.ue();
        if (s.num_ref_frames_in_pic_order_cnt_cycle > 255) return -1;
        for (int i = 0; i < s.num_ref_frames_in_pic_order_cnt_cycle; i++) s.offset_for_ref_frame[i] = br.se();
    }
    s.max_num_ref_frames = (int)br.ue();
    s.gaps_in_frame_num_value_allowed_flag = br.u1();
    s.pic_width_in_mbs_minus1 = (int)br.ue(); s.pic_height_in_map_units_minus1 = (int)br.ue();
    s.frame_mbs_only_flag = br.u1();
    if (!s.frame_mbs_only_flag) s.mb_adaptive_frame_field_flag = br.u1();
    s.direct_8x8_inference_flag = br.u1();
    if (br.u1()) { br.ue(); br.ue(); br.ue(); br.ue(); }          // frame cropping: parsed, never applied (H264PictureBase.cpp:449)
    if (br.u1()) {                                                  // vui_parameters (H264VUI.cpp:104-179)
        if (br.u1()) { if (br.u(8) == 255) { br.u(16); br.u(16); } }
        if (br.u1()) br.u1();
        if (br.u1()) { br.u(3); br.u1(); if (br.u1()) { br.u(8); br.u(8); br.u(8); } }
        if (br.u1()) { br.ue(); br.ue(); }
        if (br.u1()) { s.timing_info_present_flag = 1; s.num_units_in_tick = br.u(32); s.time_scale = br.u(32); br.u1(); }
        const int nal_hrd = br.u1(); if (nal_hrd) hrd_parameters(br);
        const int vcl_hrd = br.u1(); if (vcl_hrd) hrd_parameters(br);
        if (nal_hrd || vcl_hrd) br.u1();
        br.u1();
        if (br.u1()) { br.u1(); br.ue(); br.ue(); br.ue(); br.ue(); s.max_num_reorder_frames = (int)br.ue(); br.ue(); }
    }
    // untrusted fields: ue() returns 0xFFFFFFFF for an invalid code, which would make the sizes below 0 (division by zero in the level
    // table look-up) or overflow the macroblock count; the log2 fields are shift counts (7.4.2.1.1: 0..12)
    if ((uint32_t)s.pic_width_in_mbs_minus1 >= 1024u || (uint32_t)s.pic_height_in_map_units_minus1 >= 1024u) return -1;
    if ((uint32_t)s.log2_max_frame_num_minus4 > 12u || (uint32_t)s.log2_max_pic_order_cnt_lsb_minus4 > 12u) return -1;
    s.PicWidthInMbs = s.pic_width_in_mbs_minus1 + 1; s.PicHeightInMapUnits = s.pic_height_in_map_units_minus1 + 1;
    s.FrameHeightInMbs = (2 - s.frame_mbs_only_flag) * s.PicHeightInMapUnits;
    if (s.max_num_reorder_frames == -1) {      // H264SPS.cpp:371-397: derive from the level's MaxDpbMbs
        if ((p == 44 || p == 86 || p == 100 || p == 110 || p == 122 || p == 244) && s.constraint_set3_flag == 1) s.max_num_reorder_frames = 0;
        else {
            int maxdpb = 0;
            for (int i = 0; i < 19; i++) if (s.level_idc == kLevelMaxDpbMbs[i][0]) { int v = kLevelMaxDpbMbs[i][1] / (s.PicWidthInMbs * s.FrameHeightInMbs); maxdpb = v < 16 ? v : 16; break; }
            s.max_num_reorder_frames = maxdpb;
        }
    }
    if (s.max_num_reorder_frames > 16) return -1;
    s.ChromaArrayType = s.separate_colour_plane_flag ? 0 : s.chroma_format_idc;
    s.MaxFrameNum = 1 << (s.log2_max_frame_num_minus4 + 4);
    s.MaxPicOrderCntLsb = 1 << (s.log2_max_pic_order_cnt_lsb_minus4 + 4);
    for (int i = 0; i < s.num_ref_frames_in_pic_order_cnt_cycle; i++) s.ExpectedDeltaPerPicOrderCntCycle += s.offset_for_ref_frame[i];
    s.valid = 1;
    return 0;
}

int parse_pps(BitReader &br, PPS &p, const SPS *spss) {
    p = PPS();
    p.pps_id = (int)br.ue(); p.sps_id = (int)br.ue();
    if (p.pps_id < 0 || p.pps_id > 255 || p.sps_id < 0 || p.sps_id > 31) return -1;
    p.entropy_coding_mode_flag = br.u1(); p.bottom_field_pic_order_in_frame_present_flag = br.u1();
    p.num_slice_groups_minus1 = (int)br.ue();
    if (p.num_slice_groups_minus1 > 0) return -2;        // FMO: not supported by this front end
    p.num_ref_idx_l0_default_active_minus1 = (int)br.ue(); p.num_ref_idx_l1_default_active_minus1 = (int)br.ue();
    p.weighted_pred_flag = br.u1(); p.weighted_bipred_idc = br.u(2);
    p.pic_init_qp_minus26 = br.se(); p.pic_init_qs_minus26 = br.se(); p.chroma_qp_index_offset = br.se();
    p.deblocking_filter_control_present_flag = br.u1(); p.constrained_intra_pred_flag = br.u1(); p.redundant_pic_cnt_present_flag = br.u1();
    p.second_chroma_qp_index_offset = p.chroma_qp_index_offset;
    if (br.more_rbsp_data()) {
        p.transform_8x8_mode_flag = br.u1();
        p.pic_scaling_matrix_present_flag = br.u1();
        if (p.pic_scaling_matrix_present_flag)
            for (int i = 0; i < 6 + ((spss[p.sps_id].chroma_format_idc != 3) ? 2 : 6) * p.transform_8x8_mode_flag; i++) {
                p.pic_scaling_list_present_flag[i] = br.u1();
                if (p.pic_scaling_list_present_flag[i]) { if (i < 6) scaling_list(br, p.ScalingList4x4[i], 16, p.UseDefault4x4[i]); else scaling_list(br, p.ScalingList8x8[i - 6], 64, p.UseDefault8x8[i - 6]); }
            }
        // REF: second_chroma_qp_index_offset is only read inside the pic_scaling_matrix_present_flag block (H264PPS.cpp:207-226);
        // without a PPS scaling matrix it keeps the value of chroma_qp_index_offset whatever the stream says
        if (p.pic_scaling_matrix_present_flag) p.second_chroma_qp_index_offset = br.se();
    }
    p.valid = 1;
    return 0;
}

// Table 7-2 defaults
static const int32_t kDef4Intra[16] = {6, 13, 13, 20, 20, 20, 28, 28, 28, 28, 32, 32, 32, 37, 37, 42};
static const int32_t kDef4Inter[16] = {10, 14, 14, 20, 20, 20, 24, 24, 24, 24, 27, 27, 27, 30, 30, 34};
static const int32_t kDef8Intra[64] = {6, 10, 10, 13, 11, 13, 16, 16, 16, 16, 18, 18, 18, 18, 18, 23, 23, 23, 23, 23, 23, 25, 25, 25, 25, 25, 25, 25, 27, 27, 27, 27,
                                       27, 27, 27, 27, 29, 29, 29, 29, 29, 29, 29, 31, 31, 31, 31, 31, 31, 33, 33, 33, 33, 33, 36, 36, 36, 36, 38, 38, 38, 40, 40, 42};
static const int32_t kDef8Inter[64] = {9, 13, 13, 15, 13, 15, 17, 17, 17, 17, 19, 19, 19, 19, 19, 21, 21, 21, 21, 21, 21, 22, 22, 22, 22, 22, 22, 22, 24, 24, 24, 24,
                                       24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 27, 27, 27, 27, 27, 27, 28, 28, 28, 28, 28, 30, 30, 30, 30, 32, 32, 32, 33, 33, 35};

// H264SliceHeader.cpp:928-1160, reproduced with its deviations from Table 7-2 (the PPS branch copies transmitted lists from the
// SPS arrays, :1121/:1153; fall-back rule B keeps the SPS-derived list when the SPS carried a matrix).
static void set_scaling_lists(SliceHeader &sh) {
    const SPS &sps = sh.sps; const PPS &pps = sh.pps;
    const int n = (sps.chroma_format_idc != 3) ? 8 : 12;
    auto c4 = [&](int i, const int32_t *src) { memcpy(sh.ScalingList4x4[i], src, 64); };
    auto c8 = [&](int i, const int32_t *src) { memcpy(sh.ScalingList8x8[i], src, 256); };
    if (!sps.seq_scaling_matrix_present_flag && !pps.pic_scaling_matrix_present_flag) {
        for (int i = 0; i < n; i++) { if (i < 6) for (int k = 0; k < 16; k++) sh.ScalingList4x4[i][k] = 16; else for (int k = 0; k < 64; k++) sh.ScalingList8x8[i - 6][k] = 16; }
        return;
    }
    if (sps.seq_scaling_matrix_present_flag)
        for (int i = 0; i < n; i++) {
            if (i < 6) {
                if (!sps.seq_scaling_list_present_flag[i]) { if (i == 0) c4(i, kDef4Intra); else if (i == 3) c4(i, kDef4Inter); else c4(i, sh.ScalingList4x4[i - 1]); }
                else if (sps.UseDefault4x4[i]) c4(i, i < 3 ? kDef4Intra : kDef4Inter);
                else c4(i, sps.ScalingList4x4[i]);
            } else {
                const int j = i - 6;
                if (!sps.seq_scaling_list_present_flag[i]) { if (i == 6) c8(j, kDef8Intra); else if (i == 7) c8(j, kDef8Inter); else c8(j, sh.ScalingList8x8[j - 2]); }
                else if (sps.UseDefault8x8[j]) c8(j, (i == 6 || i == 8 || i == 10) ? kDef8Intra : kDef8Inter);
                else c8(j, sps.ScalingList8x8[j]);
            }
        }
    if (pps.pic_scaling_matrix_present_flag)
        for (int i = 0; i < n; i++) {
            if (i < 6) {
                if (!pps.pic_scaling_list_present_flag[i]) {
                    if (i == 0) { if (!sps.seq_scaling_matrix_present_flag) c4(i, kDef4Intra); }
                    else if (i == 3) { if (!sps.seq_scaling_matrix_present_flag) c4(i, kDef4Inter); }
                    else c4(i, sh.ScalingList4x4[i - 1]);
                } else if (pps.UseDefault4x4[i]) c4(i, i < 3 ? kDef4Intra : kDef4Inter);
                else c4(i, sps.ScalingList4x4[i]);
            } else {
                const int j = i - 6;
                if (!pps.pic_scaling_list_present_flag[i]) {
                    if (i == 6) { if (!sps.seq_scaling_matrix_present_flag) c8(j, kDef8Intra); }
                    else if (i == 7) { if (!sps.seq_scaling_matrix_present_flag) c8(j, kDef8Inter); }
                    else c8(j, sh.ScalingList8x8[j - 2]);
                } else if (pps.UseDefault8x8[j]) c8(j, (i == 6 || i == 8 || i == 10) ? kDef8Intra : kDef8Inter);
                else c8(j, sps.ScalingList8x8[j]);
            }
        }
}

int parse_slice_header(BitReader &br, int nal_ref_idc, int nal_unit_type, const SPS *spss, const PPS *ppss, SliceHeader &sh) {
    sh = SliceHeader();
    sh.nal_ref_idc = nal_ref_idc; sh.nal_unit_type = nal_unit_type; sh.IdrPicFlag = nal_unit_type == 5;
    sh.first_mb_in_slice = (int)br.ue();
    sh.slice_type = (int)br.ue();
    sh.pps_id = (int)br.ue();
    if (sh.slice_type < 0 || sh.slice_type > 9 || sh.pps_id < 0 || sh.pps_id > 255) return -1;
    sh.pps = ppss[sh.pps_id];
    if (!sh.pps.valid) return -1;
    sh.sps = spss[sh.pps.sps_id];
    if (!sh.sps.valid) return -1;
    const SPS &sps = sh.sps; const PPS &pps = sh.pps;
    if (sh.slice_type > 4) sh.slice_type -= 5;
    if (sps.separate_colour_plane_flag) sh.colour_plane_id = br.u(2);
    sh.frame_num = br.u(sps.log2_max_frame_num_minus4 + 4);
    if (!sps.frame_mbs_only_flag) { sh.field_pic_flag = br.u1(); if (sh.field_pic_flag) sh.bottom_field_flag = br.u1(); }
    if (sh.IdrPicFlag) sh.idr_pic_id = (int)br.ue();
    if (sps.pic_order_cnt_type == 0) {
        sh.pic_order_cnt_lsb = br.u(sps.log2_max_pic_order_cnt_lsb_minus4 + 4);
        if (pps.bottom_field_pic_order_in_frame_present_flag && !sh.field_pic_flag) sh.delta_pic_order_cnt_bottom = br.se();
    }
    if (sps.pic_order_cnt_type == 1 && !sps.delta_pic_order_always_zero_flag) {
        sh.delta_pic_order_cnt[0] = br.se();
        if (pps.bottom_field_pic_order_in_frame_present_flag && !sh.field_pic_flag) sh.delta_pic_order_cnt[1] = br.se();
    }
    if (pps.redundant_pic_cnt_present_flag) sh.redundant_pic_cnt = (int)br.ue();
    if (sh.slice_type == SLICE_B) sh.direct_spatial_mv_pred_flag = br.u1();
    if (sh.slice_type == SLICE_P || sh.slice_type == SLICE_SP || sh.slice_type == SLICE_B) {
        sh.num_ref_idx_active_override_flag = br.u1();
        sh.num_ref_idx_l0_active_minus1 = pps.num_ref_idx_l0_default_active_minus1;
        sh.num_ref_idx_l1_active_minus1 = pps.num_ref_idx_l1_default_active_minus1;
        if (sh.num_ref_idx_active_override_flag) { sh.num_ref_idx_l0_active_minus1 = (int)br.ue(); if (sh.slice_type == SLICE_B) sh.num_ref_idx_l1_active_minus1 = (int)br.ue(); }
        if ((unsigned)sh.num_ref_idx_l0_active_minus1 >= 32 || (unsigned)sh.num_ref_idx_l1_active_minus1 >= 32) return -1;
    }
    if (nal_unit_type == 20 || nal_unit_type == 21) return -1;
    for (int l = 0; l < 2; l++) {                           // ref_pic_list_modification (H264SliceHeader.cpp:516-580)
        if (l == 0 ? (sh.slice_type == SLICE_I || sh.slice_type == SLICE_SI) : (sh.slice_type != SLICE_B)) continue;
        sh.ref_pic_list_modification_flag[l] = br.u1();
        if (!sh.ref_pic_list_modification_flag[l]) continue;
        int i = 0;
        do {
            if (i >= 32) return -1;
            const int idc = (int)br.ue();
            sh.modification_of_pic_nums_idc[l][i] = idc;
            if (idc == 0 || idc == 1) sh.abs_diff_pic_num_minus1[l][i] = (int)br.ue();
            else if (idc == 2) sh.long_term_pic_num[l][i] = (int)br.ue();
            i++;
        } while (sh.modification_of_pic_nums_idc[l][i - 1] != 3);
        sh.modification_count[l] = i;
    }
    if ((pps.weighted_pred_flag && (sh.slice_type == SLICE_P || sh.slice_type == SLICE_SP)) || (pps.weighted_bipred_idc == 1 && sh.slice_type == SLICE_B)) {
        sh.luma_log2_weight_denom = (int)br.ue();                                    // pred_weight_table (H264SliceHeader.cpp:586-665)
        if (sps.ChromaArrayType != 0) sh.chroma_log2_weight_denom = (int)br.ue();
        if (sh.luma_log2_weight_denom > 30 || sh.chroma_log2_weight_denom > 30) return -1;      // (the reference has no range check; > 7 is non-conforming)
        for (int l = 0; l < (sh.slice_type == SLICE_B ? 2 : 1); l++) {
            const int n = l ? sh.num_ref_idx_l1_active_minus1 : sh.num_ref_idx_l0_active_minus1;
            for (int i = 0; i <= n; i++) {
                sh.luma_weight[l][i] = 1 << sh.luma_log2_weight_denom; sh.luma_offset[l][i] = 0;
                sh.last_luma_weight_flag[l] = br.u1();
                if (sh.last_luma_weight_flag[l]) { sh.luma_weight[l][i] = br.se(); sh.luma_offset[l][i] = br.se(); }
                if (sps.ChromaArrayType != 0) {
                    sh.chroma_weight[l][i][0] = sh.chroma_weight[l][i][1] = 1 << sh.chroma_log2_weight_denom; sh.chroma_offset[l][i][0] = sh.chroma_offset[l][i][1] = 0;
                    if (br.u1()) for (int j = 0; j < 2; j++) { sh.chroma_weight[l][i][j] = br.se(); sh.chroma_offset[l][i][j] = br.se(); }
                }
            }
        }
    }
    if (nal_ref_idc != 0) {                                  // dec_ref_pic_marking (H264SliceHeader.cpp:672-722)
        if (sh.IdrPicFlag) { sh.no_output_of_prior_pics_flag = br.u1(); sh.long_term_reference_flag = br.u1(); }
        else {
            sh.adaptive_ref_pic_marking_mode_flag = br.u1();
            if (sh.adaptive_ref_pic_marking_mode_flag) {
                int i = 0;
                do {
                    if (i >= 32) break;
                    Mmco &m = sh.mmco[i];
                    m.op = (int)br.ue();
                    if (m.op == 1 || m.op == 3) m.difference_of_pic_nums_minus1 = (int)br.ue();
                    if (m.op == 2) m.long_term_pic_num = (int)br.ue();
                    if (m.op == 3 || m.op == 6) m.long_term_frame_idx = (int)br.ue();
                    if (m.op == 4) m.max_long_term_frame_idx_plus1 = (int)br.ue();
                    i++;
                } while (sh.mmco[i - 1].op != 0);
                sh.mmco_count = i;
            }
        }
    }
    if (pps.entropy_coding_mode_flag && sh.slice_type != SLICE_I && sh.slice_type != SLICE_SI) sh.cabac_init_idc = (int)br.ue();
    sh.slice_qp_delta = br.se();
    if (sh.slice_type == SLICE_SP || sh.slice_type == SLICE_SI) { if (sh.slice_type == SLICE_SP) br.u1(); br.se(); }
    if (pps.deblocking_filter_control_present_flag) {
        sh.disable_deblocking_filter_idc = (int)br.ue();
        if (sh.disable_deblocking_filter_idc != 1) { sh.slice_alpha_c0_offset_div2 = br.se(); sh.slice_beta_offset_div2 = br.se(); }
    }
    sh.SliceQPY = 26 + pps.pic_init_qp_minus26 + sh.slice_qp_delta;
    sh.MbaffFrameFlag = (sps.mb_adaptive_frame_field_flag && !sh.field_pic_flag);
    sh.PicHeightInMbs = sps.FrameHeightInMbs / (1 + sh.field_pic_flag);
    sh.PicSizeInMbs = sps.PicWidthInMbs * sh.PicHeightInMbs;
    sh.MaxPicNum = sh.field_pic_flag ? 2 * sps.MaxFrameNum : sps.MaxFrameNum;
    sh.CurrPicNum = sh.field_pic_flag ? 2 * sh.frame_num + 1 : sh.frame_num;
    sh.FilterOffsetA = sh.slice_alpha_c0_offset_div2 * 2; sh.FilterOffsetB = sh.slice_beta_offset_div2 * 2;
    set_scaling_lists(sh);
    return 0;
}

bool first_vcl_nal_of_picture(const SliceHeader &a, const SliceHeader &b) {      // H264SliceHeader.cpp:1161-1207
    int r = 0;
    r |= a.pps_id != b.pps_id; r |= a.frame_num != b.frame_num; r |= a.field_pic_flag != b.field_pic_flag;
    if (a.field_pic_flag && b.field_pic_flag) r |= a.bottom_field_flag != b.bottom_field_flag;
    r |= (a.nal_ref_idc != b.nal_ref_idc) && (a.nal_ref_idc == 0 || b.nal_ref_idc == 0);
    r |= a.IdrPicFlag != b.IdrPicFlag;
    if (a.IdrPicFlag && b.IdrPicFlag) r |= a.idr_pic_id != b.idr_pic_id;
    if (a.sps.pic_order_cnt_type == 0) {
        r |= a.pic_order_cnt_lsb != b.pic_order_cnt_lsb;
        if (a.pps.bottom_field_pic_order_in_frame_present_flag == 1 && a.field_pic_flag == 0) r |= a.delta_pic_order_cnt_bottom != b.delta_pic_order_cnt_bottom;
    }
    if (a.sps.pic_order_cnt_type == 1 && !a.sps.delta_pic_order_always_zero_flag) {
        r |= a.delta_pic_order_cnt[0] != b.delta_pic_order_cnt[0];
        if (a.pps.bottom_field_pic_order_in_frame_present_flag == 1 && a.field_pic_flag == 0) r |= a.delta_pic_order_cnt[1] != b.delta_pic_order_cnt[1];
    }
    return r != 0;
}

// ------------------------------------------------------------------ CABAC
void Cabac::init_contexts(int slice_type, int cabac_init_idc, int slice_qp) {
    const int t = (slice_type == SLICE_I || slice_type == SLICE_SI) ? 0 : 1 + cabac_init_idc;
    const int qp = slice_qp < 0 ? 0 : slice_qp > 51 ? 51 : slice_qp;
    for (int i = 0; i < 1024; i++) {
        int pre = ((kCabacMN[t][i][0] * qp) >> 4) + kCabacMN[t][i][1];
        pre = pre < 1 ? 1 : pre > 126 ? 126 : pre;
        state[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
    }
}

const uint8_t (&g_range_lps)[64][4] = kRangeTabLPS;
uint8_t g_next_state[2][128];
static const bool g_next_state_init = [] {
    for (int s = 0; s < 128; s++) {
        const int p = s >> 1, mps = s & 1;
        g_next_state[0][s] = (uint8_t)((kTransIdxMPS[p] << 1) | mps);
        g_next_state[1][s] = (uint8_t)((kTransIdxLPS[p] << 1) | (p == 0 ? mps ^ 1 : mps));
    }
    return true;
}();

const uint16_t (*coeff_token_table())[17][4][2] { return kCoeffToken; }
const uint16_t (*total_zeros_table())[16][16][2] { return kTotalZeros; }
const uint16_t (*run_before_table())[15][2] { return kRunBefore; }
const uint8_t (*me_cbp_table())[2] { return kMeCbp; }
const uint8_t *scan4x4_table(int field) { return kScan4x4[field]; }
const uint8_t *scan8x8_table(int field) { return kScan8x8[field]; }

}  // namespace h264b2
