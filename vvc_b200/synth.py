"""Seeded synthetic pictures WITH synthetic side information for the in-loop filter chain.

The kernels are pure functions of (planes, side info), so a picture does not have to come out of a
decoder to exercise them: this module draws a random quadtree block structure and derives from it deblocking
segment records that obey the same geometric rules the reference's derivation guarantees (filter lengths from the
block sizes on both sides of an edge, LoopFilter.cpp:482-497 -- that is what makes the edges of one pass
independent of each other), plus random SAO CTU parameters and random ALF / CC-ALF APS data and CTU controls.
`density` scales how much of each tool is switched on; density=1.0 is the "forced-on" stress picture.

Used by the parity tests (GPU vs oracle, any size / chroma format) and by bench.py when no captured pictures exist.
"""
import ctypes as C

import numpy as np

from . import abi
from .capture import Capture

# tc / beta of the standard for 10-bit content at a few plausible QPs (LoopFilter.cpp:66-74)
_TC = [0, 3, 4, 5, 7, 9, 11, 14, 19, 25, 33, 45, 64, 89]
_BETA = [0, 24, 32, 40, 48, 56, 64, 88, 104, 128, 160, 200, 256, 352]


def _quadtree(rng, w, h, ctu, p_split):
    """Per-4x4-unit log2 block size (2..log2 ctu) of a random quadtree partition aligned to the CTU grid."""
    uh, uw = h // 4, w // 4
    size = np.full((uh, uw), int(np.log2(ctu)), dtype=np.int8)
    lvl = int(np.log2(ctu))
    while lvl > 2:
        bu = (1 << lvl) // 4                                  # block edge in units
        nby, nbx = (uh + bu - 1) // bu, (uw + bu - 1) // bu
        split = rng.random((nby, nbx)) < p_split
        split = np.repeat(np.repeat(split, bu, axis=0), bu, axis=1)[:uh, :uw]
        size = np.where((size == lvl) & split, lvl - 1, size).astype(np.int8)
        lvl -= 1
    return size


def _records(rng, size, w, h, ctu, sx, sy, density, bd):
    uh, uw = h // 4, w // 4
    ys, xs = np.mgrid[0:uh, 0:uw]
    blk = (1 << size.astype(np.int32))                         # block edge in luma samples at each unit
    luma, chroma = [], []
    shift = bd - 10 if bd >= 10 else 0
    for d in range(2):
        pos = (xs if d == 0 else ys) * 4
        edge = (pos % blk) == 0
        edge &= pos > 0
        sizeQ = blk
        sizeP = np.roll(blk, 1, axis=1 if d == 0 else 0)
        small = (sizeP <= 4) | (sizeQ <= 4)
        lenP = np.where(small, 1, np.where(sizeP >= 32, 7, 3))
        lenQ = np.where(small, 1, np.where(sizeQ >= 32, 7, 3))
        # a few "sub-block" style lengths (2 and 5) where they are legal: blocks of at least 16 / 32 samples
        lenP = np.where((sizeP >= 16) & (sizeQ >= 16) & (rng.random(edge.shape) < 0.05), 2, lenP)
        lenQ = np.where(lenP == 2, 2, lenQ)
        lenP = np.where((lenP == 7) & (rng.random(edge.shape) < 0.2), 5, lenP)
        lenQ = np.where((lenQ == 7) & (rng.random(edge.shape) < 0.2), 5, lenQ)
        on = edge & (rng.random(edge.shape) < density)
        qi = rng.integers(1, len(_TC), size=edge.shape)
        tc = np.array(_TC)[qi] << shift
        beta = (np.array(_BETA)[np.minimum(len(_BETA) - 1, qi + rng.integers(-1, 2, size=edge.shape)).clip(0)] << shift).clip(0, 0x7FF)
        rec = tc.astype(np.uint32) | (beta.astype(np.uint32) << 11) | (lenP.astype(np.uint32) << 22) | (lenQ.astype(np.uint32) << 25)
        rec |= np.where(rng.random(edge.shape) < 0.02, np.uint32(abi.DBF_L_PNOFILT), np.uint32(0))
        rec |= np.where(rng.random(edge.shape) < 0.02, np.uint32(abi.DBF_L_QNOFILT), np.uint32(0))
        if d == 1:
            rec |= np.where((pos % ctu) == 0, np.uint32(abi.DBF_L_CTUROW), np.uint32(0))
        luma.append(np.where(on, rec, 0).astype(np.uint32).ravel())

        # chroma: edges on the 8x8 chroma-sample grid; large iff both chroma blocks are >= 8 samples across the edge
        s_across = sx if d == 0 else sy
        g = 8 << s_across
        cP, cQ = sizeP >> s_across, sizeQ >> s_across
        large = (cP >= 8) & (cQ >= 8)
        con = edge & ((pos % g) == 0) & (rng.random(edge.shape) < density)
        crec = np.zeros(edge.shape, dtype=np.uint64)
        for c in range(2):
            qc = rng.integers(1, len(_TC), size=edge.shape)
            tcc = (np.array(_TC)[qc] << shift).astype(np.uint64)
            bc = (np.array(_BETA)[qc] << shift).clip(0, 0x7FF).astype(np.uint64)
            use = rng.random(edge.shape) < 0.85
            crec |= np.where(use, tcc, 0).astype(np.uint64) << np.uint64(11 * c)
            crec |= np.where(use & large, bc, 0).astype(np.uint64) << np.uint64(22 + 11 * c)
        crec |= np.where(large, np.uint64(abi.DBF_C_LARGE), np.uint64(0))
        if d == 1:
            crec |= np.where((pos % ctu) == 0, np.uint64(abi.DBF_C_CTB), np.uint64(0))
        crec |= np.where(rng.random(edge.shape) < 0.02, np.uint64(abi.DBF_C_PNOFILT), np.uint64(0))
        crec = np.where(con & ((crec & np.uint64(0x3FFFFF)) != 0), crec, np.uint64(0))
        if d == 0:
            cols = (w + g - 1) // g
            out = np.zeros((uh, cols), dtype=np.uint64)
            sel = np.arange(0, uw, g // 4)
            out[:, :len(sel)] = crec[:, sel]
        else:
            rows = (h + g - 1) // g
            out = np.zeros((rows, uw), dtype=np.uint64)
            sel = np.arange(0, uh, g // 4)
            out[:len(sel), :] = crec[sel, :]
        chroma.append(out.ravel())
    return luma, chroma


def _sao(rng, nctus, wctus, hctus, ncomp, density, bd):
    arr = (abi.SaoCtu * nctus)()
    maxoff = (1 << (min(bd, 10) - 5)) - 1
    for a in range(nctus):
        cx, cy = a % wctus, a // wctus
        avail = 0
        for bit, (dx, dy) in zip((1, 2, 4, 8, 16, 32, 64, 128), ((-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (1, -1), (-1, 1), (1, 1))):
            inside = 0 <= cx + dx < wctus and 0 <= cy + dy < hctus
            if inside and rng.random() > 0.08:       # a few "other slice" neighbours
                avail |= bit
        arr[a].avail = avail
        arr[a].merge_left_ok = int(cx > 0)
        arr[a].merge_above_ok = int(cy > 0)
        for c in range(ncomp):
            o = arr[a].comp[c]
            r = rng.random()
            if r > density:
                o.mode = abi.SAO_MODE_OFF
                continue
            if r < 0.25 * density and (cx > 0 or cy > 0):
                o.mode = abi.SAO_MODE_MERGE
                o.type = abi.SAO_MERGE_LEFT if (cx > 0 and (cy == 0 or rng.random() < 0.5)) else abi.SAO_MERGE_ABOVE
                continue
            o.mode = abi.SAO_MODE_NEW
            o.type = int(rng.integers(0, 5))
            if o.type == abi.SAO_BO:
                o.aux = int(rng.integers(0, 32))
                for i in range(4):
                    o.offset[(o.aux + i) & 31] = int(rng.integers(-maxoff, maxoff + 1))
            else:
                v = rng.integers(0, maxoff + 1, size=4)
                o.offset[0], o.offset[1], o.offset[2], o.offset[3], o.offset[4] = int(v[0]), int(v[1]), 0, -int(v[2]), -int(v[3])
    return arr


def _alf_sections(rng, nctus, density, chroma):
    naps = int(rng.integers(1, 4))
    luma = (abi.AlfLumaAps * naps)()
    for s in range(naps):
        a = luma[s]
        a.num_filters = int(rng.integers(1, 26))
        a.nonlinear = int(rng.random() < 0.7)
        for c in range(25):
            a.delta_idx[c] = int(rng.integers(0, a.num_filters))
        for f in range(25):
            v = rng.integers(-24, 25, size=12)
            for k in range(12):
                a.coeff[f][k] = int(v[k])
                a.clip_idx[f][k] = int(rng.integers(0, 4))
    ch = abi.AlfChromaAps()
    ch.num_alts = int(rng.integers(1, 9))
    ch.nonlinear = int(rng.random() < 0.7)
    for alt in range(8):
        v = rng.integers(-30, 31, size=6)
        for k in range(6):
            ch.coeff[alt][k] = int(v[k])
            ch.clip_idx[alt][k] = int(rng.integers(0, 4))
    en = [(rng.random(nctus) < density).astype(np.uint8) for _ in range(3)]
    fidx = rng.integers(0, 16 + naps, size=nctus).astype(np.int16)
    alt = [rng.integers(0, ch.num_alts, size=nctus).astype(np.uint8) for _ in range(2)]
    cc_en = [int(chroma and rng.random() < max(density, 0.5)) for _ in range(2)]
    cc_coef = (rng.integers(-3, 4, size=(2, 4, 8)) * (1 << rng.integers(0, 4, size=(2, 4, 8)))).astype(np.int16)
    cc_idc = [np.where(rng.random(nctus) < density, rng.integers(1, 5, size=nctus), 0).astype(np.uint8) for _ in range(2)]
    hdr = np.array([1, int(chroma), int(chroma), naps, int(chroma), cc_en[0], cc_en[1], nctus], dtype=np.int32)
    if not chroma:
        en[1][:] = 0
        en[2][:] = 0
    sec = {"alf_hdr": hdr.tobytes(), "alf_luma_aps": bytes(luma), "alf_chroma_aps": bytes(ch), "alf_fidx": fidx.tobytes(),
           "alf_cccoef": cc_coef.tobytes()}
    for c in range(3):
        sec["alf_en%d" % c] = en[c].tobytes()
    for c in range(2):
        sec["alf_alt%d" % c] = alt[c].tobytes()
        sec["alf_ccidc%d" % c] = cc_idc[c].tobytes()
    return sec


def _plane(rng, h, w, bd, phase):
    yy, xx = np.mgrid[0:h, 0:w]
    mid, amp = 1 << (bd - 1), (1 << bd) / 4.0
    p = (mid + amp * np.sin((xx + 13 * phase) / 37.0) * np.cos((yy - 7 * phase) / 23.0)
         + 0.5 * amp * (((xx // 16) + (yy // 16) + phase) % 2) + rng.normal(0, amp / 12.0, (h, w)))
    # saturated patches exercise the clipping paths
    p[: h // 16, : w // 8] = (1 << bd) - 1
    p[-(h // 16):, -(w // 8):] = 0
    return np.clip(np.rint(p), 0, (1 << bd) - 1).astype(np.int16)


def make_picture(width, height, chroma_format=1, bit_depth=10, ctu_size=128, seed=0, density=0.6, p_split=0.55,
                 dbf=True, sao=True, alf=True, partitions=False, ladf=False, vb=None):
    """Returns a Capture (pre planes + side info, no reference stage outputs).  partitions: random per-CTU ALF clip / corner-pad
    flags as slice and tile boundaries without cross-boundary filtering produce them (any combination is a legal input).
    ladf: the luma deblocking records carry QPs and the sequence has LADF intervals (include/vtmgpu.h, vtmgpu_ladf).
    vb: ([x positions], [y positions]) of signalled virtual boundaries (multiples of 8, at least a CTU apart), seen by SAO and ALF."""
    assert width % 8 == 0 and height % 8 == 0
    rng = np.random.default_rng(seed)
    sx, sy = abi.chroma_shifts(chroma_format)
    ncomp = 1 if chroma_format == 0 else 3
    wctus, hctus = (width + ctu_size - 1) // ctu_size, (height + ctu_size - 1) // ctu_size
    nctus = wctus * hctus
    sec = {"seq": np.array([width, height, chroma_format, bit_depth, bit_depth, ctu_size, seed, 0], dtype=np.int32).tobytes()}
    sec["pre_0"] = _plane(rng, height, width, bit_depth, seed).tobytes()
    if ncomp == 3:
        for c in (1, 2):
            sec["pre_%d" % c] = _plane(rng, height >> sy, width >> sx, bit_depth, seed + 3 * c).tobytes()
    size = _quadtree(rng, width, height, ctu_size, p_split)
    luma, chroma = _records(rng, size, width, height, ctu_size, sx, sy, density if dbf else 0.0, bit_depth)
    if ladf:
        lrng = np.random.default_rng(seed + 555)
        for d in range(2):
            n = luma[d].size
            qp_tc = lrng.integers(8, 70, size=n).astype(np.uint32)
            qp_b = (qp_tc.astype(np.int64) + lrng.integers(-6, 7, size=n)).clip(0, 80).astype(np.uint32)
            rec = (luma[d] & np.uint32(~0x3FFFFF & 0xFFFFFFFF)) | (qp_tc + abi.DBF_LADF_BIAS) | ((qp_b + abi.DBF_LADF_BIAS) << 11)
            luma[d] = np.where(luma[d] != 0, rec, 0).astype(np.uint32)
        full = 1 << bit_depth
        nint = int(lrng.integers(2, 6))
        offs = [int(v) for v in lrng.integers(-6, 7, size=5)]
        lbs = [0] + sorted(int(v) for v in lrng.integers(full // 8, full - full // 8, size=4))
        sec["dbf_ladf"] = np.array([nint] + offs + lbs, dtype=np.int32).tobytes()
    for d in range(2):
        sec["dbfrec_l%d" % d] = luma[d].tobytes()
        sec["dbfrec_c%d" % d] = chroma[d].tobytes() if ncomp == 3 else b""
    if vb is not None:
        vx, vy = list(vb[0]), list(vb[1])
        sec["vb"] = np.array([len(vx), len(vy)] + (vx + [0, 0, 0])[:3] + (vy + [0, 0, 0])[:3], dtype=np.int32).tobytes()
    if sao:
        sec["sao_raw"] = bytes(_sao(rng, nctus, wctus, hctus, ncomp, density, bit_depth))
        sec["sao_scale"] = np.array([0, 0], dtype=np.int32).tobytes()
    if alf:
        sec.update(_alf_sections(rng, nctus, density, ncomp == 3))
        if partitions:
            prng = np.random.default_rng(seed + 777)       # separate stream: the other sections do not depend on this option
            clip = np.zeros(nctus, dtype=np.uint8)
            for a in range(nctus):
                cx, cy = a % wctus, a // wctus
                f = int(prng.integers(0, 16)) if prng.random() < 0.7 else 0
                if cy == 0: f &= ~abi.ALF_CLIP_TOP
                if cy == hctus - 1: f &= ~abi.ALF_CLIP_BOTTOM
                if cx == 0: f &= ~abi.ALF_CLIP_LEFT
                if cx == wctus - 1: f &= ~abi.ALF_CLIP_RIGHT
                if not f & (abi.ALF_CLIP_TOP | abi.ALF_CLIP_LEFT) and cx > 0 and cy > 0 and prng.random() < 0.5:
                    f |= abi.ALF_PAD_TL
                if not f & (abi.ALF_CLIP_BOTTOM | abi.ALF_CLIP_RIGHT) and cx < wctus - 1 and cy < hctus - 1 and prng.random() < 0.5:
                    f |= abi.ALF_PAD_BR
                clip[a] = f
            sec["alf_clip"] = clip.tobytes()
    return Capture(sec)


def make_alf_slices(cap, nslices=3, seed=0, off_slice=None):
    """A multi-slice ALF description for the picture of `cap` (vtmgpu_set_alf_slices): slice 0 carries the capture's own
    parameters, slices 1.. freshly drawn ones (other APSs, chroma alternatives, enable flags); off_slice = index of a slice that
    switches ALF off altogether.  The slices are horizontal runs of CTUs in raster order (cut at random CTUs), the per-CTU arrays of
    every CTU are drawn for its own slice.  Returns ([abi.AlfParams per slice], ctu_slice uint8[ctus]); the per-picture arrays live
    in the first element, as the entry point reads them."""
    import ctypes as C
    rng = np.random.default_rng(1000 + seed)
    n = cap.num_ctus
    cuts = sorted(rng.choice(np.arange(1, n), size=nslices - 1, replace=False).tolist()) if nslices > 1 else []
    ctu_slice = np.zeros(n, dtype=np.uint8)
    for k, c0 in enumerate(cuts):
        ctu_slice[c0:] = k + 1
    chroma = cap.ncomp == 3
    secs = [None] + [_alf_sections(rng, n, 0.9, chroma) for _ in range(nslices - 1)]
    slices, keep = [], []
    base = cap.alf
    comb = dict(ctu_enable=[a.copy() for a in base["ctu_enable"]], filter_idx=base["filter_idx"].copy(), ctu_alt=[a.copy() for a in base["ctu_alt"]],
                cc_idc=[a.copy() for a in base["cc_idc"]])
    for k in range(nslices):
        if k == 0:
            p = cap.alf_params()
        else:
            sec = secs[k]
            hdr = np.frombuffer(sec["alf_hdr"], dtype=np.int32)
            p = abi.AlfParams()
            for c in range(3):
                p.enabled[c] = int(hdr[c])
            p.num_luma_aps = min(int(hdr[3]), 2 if nslices <= 3 else 1)           # a picture holds at most 8 distinct luma APSs
            la = (abi.AlfLumaAps * int(hdr[3])).from_buffer_copy(sec["alf_luma_aps"])
            ca = abi.AlfChromaAps.from_buffer_copy(sec["alf_chroma_aps"])
            keep += [la, ca]
            p.luma_aps = C.cast(la, C.POINTER(abi.AlfLumaAps))
            if chroma:
                p.chroma_aps = C.pointer(ca)
            for c in range(2):
                p.ccalf_enabled[c] = int(hdr[5 + c])
            p.num_ctus = n
            m = ctu_slice == k
            for c in range(3):
                comb["ctu_enable"][c][m] = np.frombuffer(sec["alf_en%d" % c], dtype=np.uint8)[m]
            comb["filter_idx"][m] = np.frombuffer(sec["alf_fidx"], dtype=np.int16)[m] % (16 + p.num_luma_aps)
            for c in range(2):
                comb["ctu_alt"][c][m] = np.frombuffer(sec["alf_alt%d" % c], dtype=np.uint8)[m]
                comb["cc_idc"][c][m] = np.frombuffer(sec["alf_ccidc%d" % c], dtype=np.uint8)[m]
        if off_slice is not None and k == off_slice:
            for c in range(3):
                p.enabled[c] = 0
        slices.append(p)
    p0 = slices[0]
    for c in range(3):
        p0.ctu_enable[c] = comb["ctu_enable"][c].ctypes.data_as(C.POINTER(C.c_uint8))
    p0.ctu_filter_idx = comb["filter_idx"].ctypes.data_as(C.POINTER(C.c_int16))
    for c in range(2):
        p0.ctu_alt[c] = comb["ctu_alt"][c].ctypes.data_as(C.POINTER(C.c_uint8))
        p0.ccalf_idc[c] = comb["cc_idc"][c].ctypes.data_as(C.POINTER(C.c_uint8))
    p0._keep2 = (comb, keep)
    return slices, ctu_slice
