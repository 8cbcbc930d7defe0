"""Seeded synthetic pictures in the engine's SoA format (TEST INFRASTRUCTURE).

The generator produces *arbitrary but well-formed* inputs — every macroblock class, every intra mode (also
ones whose neighbours are missing, Q15), 16x16..4x4 motion with vectors far outside the picture, field
views, explicit weights, custom scaling lists, slices, all three deblocking idc values, trailing
never-decoded macroblocks, MBAFF field/frame pairs — so the CUDA path and the CPU oracle can be compared
bit for bit on cases the bundled streams never reach.
"""
import numpy as np

from h264_video_decoder_demo_b200 import abi
from h264_video_decoder_demo_b200.replay import Picture, Replay


def _levels(rng, n, amp, density):
    v = rng.integers(-amp, amp + 1, size=n)
    v[rng.random(n) > density] = 0
    return v.astype(np.int16)


def synth_picture(rng, wmb, hmb, *, inter_frac=0.6, mbaff=False, dst=0, ref_slots=(1, 2), n_slices=3, na_tail=0,
                  custom_scaling=False, pcm_frac=0.02, cip=False, amp=60, decode_idx=0, deblock=True):
    nmb = wmb * hmb
    info = np.zeros(nmb, dtype=abi.MB_INFO_DT)
    modes = np.zeros(nmb, dtype="<u8")
    coff = np.zeros(nmb, dtype="<u4")
    motion = np.zeros(nmb, dtype=abi.MB_MOTION_DT)
    motion["ref_surf"] = -1
    motion["ref_ident"] = -1
    n_w = 6
    weights = np.zeros(n_w, dtype=abi.WEIGHT_DT)
    weights["w0"][0] = 1
    weights["w1"][0] = 1
    for i in range(1, n_w):
        weights["mode"][i] = 1
        weights["logwd"][i] = rng.integers(0, 8, 3)
        weights["w0"][i] = rng.integers(-40, 100, 3)
        weights["w1"][i] = rng.integers(-40, 100, 3)
        weights["o0"][i] = rng.integers(-20, 21, 3)
        weights["o1"][i] = rng.integers(-20, 21, 3)
    coefs = []
    ncoef = 0
    bounds = sorted(rng.choice(np.arange(1, nmb), size=min(n_slices - 1, nmb - 1), replace=False).tolist()) if n_slices > 1 else []
    if mbaff:
        bounds = [b & ~1 for b in bounds]
    slice_of = np.searchsorted(np.array(bounds + [nmb]), np.arange(nmb), side="right")
    slice_idc = rng.integers(0, 3, n_slices + 1)
    slice_oa = rng.integers(-3, 4, n_slices + 1) * 2
    slice_ob = rng.integers(-3, 4, n_slices + 1) * 2
    field_pair = rng.random((nmb + 1) // 2) < 0.4 if mbaff else None
    W4 = wmb * 64
    has_inter = 0
    first_na = nmb - na_tail
    for a in range(nmb):
        I = info[a]
        coff[a] = ncoef
        if a >= first_na:
            I["mb_class"] = abi.MB_NA
            continue
        r = rng.random()
        if r < inter_frac:
            cls = abi.MB_INTER
        elif r < inter_frac + pcm_frac:
            cls = abi.MB_IPCM
        else:
            cls = int(rng.choice([abi.MB_I4x4, abi.MB_I8x8, abi.MB_I16x16]))
        field = bool(mbaff and field_pair[a >> 1])
        t8 = cls == abi.MB_I8x8 or (cls == abi.MB_INTER and rng.random() < 0.4) or (cls == abi.MB_I16x16 and rng.random() < 0.1)
        flags = (abi.MBF_FIELD if field else 0) | (abi.MBF_T8x8 if t8 else 0)
        if rng.random() < 0.03:
            flags |= abi.MBF_SPSI
        if cip and cls == abi.MB_INTER and rng.random() < 0.5:
            flags |= abi.MBF_CIP_UNAVAIL
        I["mb_class"] = cls
        I["flags"] = flags
        I["pred16_chroma"] = int(rng.integers(0, 4)) | (int(rng.integers(0, 4)) << 2)
        I["qpy"] = int(rng.integers(0, 52))
        s = int(slice_of[a])
        I["slice_number"] = s
        I["nnz_mask"] = int(rng.integers(0, 1 << 16)) if rng.random() < 0.5 else 0
        I["filter_offset_a"] = int(slice_oa[s])
        I["filter_offset_b"] = int(slice_ob[s])
        I["deblock_idc"] = int(slice_idc[s])
        cm = 0
        if cls == abi.MB_IPCM:
            cm = abi.CM_PCM
            coefs.append(rng.integers(0, 256, 384).astype(np.int16))
            ncoef += 384
        else:
            dens = rng.choice([0.0, 0.1, 0.5])
            eff_t8 = t8 and cls != abi.MB_I16x16
            for b in range(4 if eff_t8 else 16):
                if rng.random() < 0.5 and dens > 0:
                    cm |= abi.CM_LUMA(b)
                    blk = _levels(rng, 64 if eff_t8 else 16, amp, dens)
                    if cls == abi.MB_I16x16:
                        blk[0] = 0
                    coefs.append(blk)
                    ncoef += blk.size
            if cls == abi.MB_I16x16 and rng.random() < 0.7:
                cm |= abi.CM_LUMA_DC
                coefs.append(_levels(rng, 16, amp * 4, 0.6))
                ncoef += 16
            if rng.random() < 0.5:
                cm |= abi.CM_CHROMA_DC
                coefs.append(_levels(rng, 8, amp * 2, 0.7))
                ncoef += 8
            for c in range(2):
                for b in range(4):
                    if rng.random() < 0.3:
                        cm |= abi.CM_CB(b) if c == 0 else abi.CM_CR(b)
                        blk = _levels(rng, 16, amp, 0.4)
                        blk[0] = 0
                        coefs.append(blk)
                        ncoef += 16
        I["coef_mask"] = cm
        if cls in (abi.MB_I4x4, abi.MB_I8x8):
            m = 0
            for b in range(16 if cls == abi.MB_I4x4 else 4):
                m |= int(rng.integers(0, 9)) << (4 * b)
            modes[a] = m
        if cls == abi.MB_INTER:
            has_inter = 1
            M = motion[a]
            shape = rng.integers(0, 5)   # 0:16x16 1:16x8 2:8x16 3:8x8 4:4x4
            big = rng.random() < 0.15
            def mv():
                if big:
                    return rng.integers(-W4 - 200, W4 + 200, 2)
                return rng.integers(-40, 41, 2)
            base = [[mv() for _ in range(2)] for _ in range(16)]
            for l in range(2):
                for r4 in range(16):
                    bx, by = r4 & 3, r4 >> 2
                    if shape == 0:
                        src = 0
                    elif shape == 1:
                        src = (by >> 1) * 8
                    elif shape == 2:
                        src = (bx >> 1) * 2
                    elif shape == 3:
                        src = (by >> 1) * 8 + (bx >> 1) * 2
                    else:
                        src = r4
                    M["mv"][l][r4] = base[src][l]
            for q in range(4):
                qq = q if shape >= 3 else (q & 2 if shape == 1 else (q & 1 if shape == 2 else 0))
                if q != qq:
                    for l in range(2):
                        M["ref_surf"][l][q] = M["ref_surf"][l][qq]
                        M["ref_ident"][l][q] = M["ref_ident"][l][qq]
                    M["wt_idx"][q] = M["wt_idx"][qq]
                    continue
                use = rng.integers(1, 4)     # bit0 L0, bit1 L1
                for l in range(2):
                    if use & (1 << l):
                        slot = int(rng.choice(ref_slots))
                        view = int(rng.integers(1, 3)) if field else 0
                        M["ref_surf"][l][q] = (slot << 2) | view
                    if rng.random() < 0.9:
                        M["ref_ident"][l][q] = int(rng.integers(-1, 4))
                M["wt_idx"][q] = int(rng.integers(0, n_w)) if rng.random() < 0.5 else 0
    coef_arr = np.concatenate(coefs).astype(np.int16) if coefs else np.zeros(0, np.int16)
    pad = (-coef_arr.size) % 4
    if pad:
        coef_arr = np.concatenate([coef_arr, np.zeros(pad, np.int16)])
    ls4 = ls8 = None
    if custom_scaling:
        ls4 = rng.integers(16, 255 * 29, 2 * 2 * 6 * 16).astype(np.int16)
        ls8 = rng.integers(16, 255 * 58 // 2, 2 * 2 * 6 * 64).astype(np.int16)
    return Picture(decode_idx=decode_idx, dst_surface=dst, clear_surface=1 if na_tail else 0, has_inter=has_inter,
                   deblock_enable=1 if deblock else 0, deblock_stop_mb=first_na, mbaff=1 if mbaff else 0,
                   cqp=(int(rng.integers(-12, 13)), int(rng.integers(-12, 13))), slice_type=0, poc=0, n_na=na_tail, nal_ref_idc=1,
                   sum_pre=0, sum_post=0, mb_info=info, intra_modes=modes, coef_offset=coff,
                   motion=motion if has_inter else None, weights=weights, coefs=coef_arr, level_scale4=ls4, level_scale8=ls8)


def synth_replay(wmb, hmb):
    return Replay("<synthetic>", wmb, hmb)


def random_surface(rng, wmb, hmb, smooth=False):
    n = wmb * hmb * 384
    if not smooth:
        return rng.integers(0, 256, n).astype(np.uint8)
    base = rng.integers(0, 256)
    return np.clip(base + rng.integers(-6, 7, n), 0, 255).astype(np.uint8)
