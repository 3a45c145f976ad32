"""oracle/pyoracle.py -- TEST INFRASTRUCTURE: ctypes front end of oracle/_build/libvvcoracle.so, the plain-C
restatement of the reference filter arithmetic (oracle/vvc_filters_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module --
as the checker or the CPU baseline, never on the product path.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from vvc_b200 import abi  # noqa: E402

LIB_PATH = os.path.join(HERE, "_build", "libvvcoracle.so")


def build(force=False):
    src = os.path.join(HERE, "vvc_filters_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.run(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(HERE, "..", "include"),
                        "-o", LIB_PATH, src], check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        common = [abi.PlanePtrs, abi.Strides, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.vvco_deblock.argtypes = common + [C.POINTER(abi.DeblockParams)]
        _lib.vvco_sao.argtypes = common + [C.c_int, C.POINTER(abi.SaoParams)]
        _lib.vvco_alf.argtypes = common + [C.c_int, C.POINTER(abi.AlfParams)]
        _lib.vvco_sao_reconstruct.argtypes = [C.POINTER(abi.SaoCtu), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.vvco_filter_picture.argtypes = common + [C.c_int, C.POINTER(abi.DeblockParams), C.POINTER(abi.SaoParams), C.POINTER(abi.AlfParams)]
    return _lib


def _planes(planes):
    ptrs, strides = abi.PlanePtrs(), abi.Strides()
    for c, p in enumerate(planes):
        assert p.dtype == np.int16 and p.flags["C_CONTIGUOUS"]
        ptrs[c] = p.ctypes.data_as(C.POINTER(C.c_int16))
        strides[c] = p.shape[1]
    return ptrs, strides


def _seq(seq):
    return [seq["width"], seq["height"], seq["chroma_format"], seq["bit_depth_luma"], seq["bit_depth_chroma"]]


def deblock(seq, planes, dbf_params):
    """In place on the list of int16 planes."""
    ptrs, strides = _planes(planes)
    rc = lib().vvco_deblock(ptrs, strides, *_seq(seq), C.byref(dbf_params))
    if rc:
        raise RuntimeError("vvco_deblock rc=%d" % rc)


def sao_reconstruct(ctus, width_in_ctus, ncomp, scale_luma, scale_chroma):
    rc = lib().vvco_sao_reconstruct(ctus, len(ctus), width_in_ctus, ncomp, scale_luma, scale_chroma)
    if rc < 0:
        raise RuntimeError("vvco_sao_reconstruct rc=%d" % rc)
    return rc


def sao(seq, planes, ctus, vb=None):
    """ctus: RECONSTRUCTED (abi.SaoCtu * n); vb: abi.VirtualBoundaries or None."""
    ptrs, strides = _planes(planes)
    p = abi.SaoParams(C.cast(ctus, C.POINTER(abi.SaoCtu)), len(ctus), C.pointer(vb) if vb is not None else None)
    rc = lib().vvco_sao(ptrs, strides, *_seq(seq), seq["ctu_size"], C.byref(p))
    if rc:
        raise RuntimeError("vvco_sao rc=%d" % rc)


def alf(seq, planes, alf_params):
    ptrs, strides = _planes(planes)
    rc = lib().vvco_alf(ptrs, strides, *_seq(seq), seq["ctu_size"], C.byref(alf_params))
    if rc:
        raise RuntimeError("vvco_alf rc=%d" % rc)


def lmcs_inverse(planes, inv_lut):
    """LMCS inverse mapping of the luma plane, in place: dst[x] = pLUT[src[x]] over the whole plane (AreaBuf<Pel>::rspSignal,
    CommonLib/Buffer.cpp:380-393, as DecLib::executeLoopFilters calls it before the deblocking, DecLib.cpp:570-577)."""
    lut = np.asarray(inv_lut, dtype=np.int16)
    planes[0][...] = lut[planes[0].astype(np.int64)]


def extend_border(seq, planes, margin_luma):
    """Picture::extendPicBorder (CommonLib/Picture.cpp:737-772, no wrap-around): left / right margins replicate the first / last sample
    of each row, then the first / last padded row is copied upwards / downwards.  Returns the padded planes."""
    sx, sy = abi.chroma_shifts(seq["chroma_format"])
    out = []
    for c, p in enumerate(planes):
        xm, ym = (margin_luma >> sx, margin_luma >> sy) if c else (margin_luma, margin_luma)
        h, w = p.shape
        q = np.empty((h + 2 * ym, w + 2 * xm), dtype=np.int16)
        q[ym:ym + h, xm:xm + w] = p
        q[ym:ym + h, :xm] = p[:, :1]
        q[ym:ym + h, xm + w:] = p[:, -1:]
        q[:ym, :] = q[ym:ym + 1, :]
        q[ym + h:, :] = q[ym + h - 1:ym + h, :]
        out.append(q)
    return out


def alf_slices(seq, planes, slices, ctu_slice):
    """ALF of a picture whose slices carry different parameters, in place.  ALFProcess reloads the APS data whenever the CTU's
    slice changes and reads the slice's own enable flags (AdaptiveLoopFilter.cpp:429-441, :451, :532); the filters read only the
    pre-ALF picture, so the result is the per-slice run of the single-set oracle on the slice's own CTUs (every other CTU switched
    off), composed CTU by CTU.  The per-CTU arrays and the CC-ALF coefficients are per picture: taken from slices[0]."""
    n = slices[0].num_ctus
    ctu = seq["ctu_size"]
    wc = (seq["width"] + ctu - 1) // ctu
    sx, sy = abi.chroma_shifts(seq["chroma_format"])
    src = [p.copy() for p in planes]
    p0 = slices[0]

    def arr(ptr, dt):
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt).copy() if ptr else None

    en = [arr(p0.ctu_enable[c], np.uint8) for c in range(3)]
    idc = [arr(p0.ccalf_idc[c], np.uint8) for c in range(2)]
    for k, q in enumerate(slices):
        mine = np.asarray(ctu_slice) == k
        if not mine.any():
            continue
        r = abi.AlfParams()
        C.memmove(C.byref(r), C.byref(p0), C.sizeof(abi.AlfParams))         # per-picture members (arrays, clip flags, vb, CC-ALF coefficients)
        for c in range(3):
            r.enabled[c] = q.enabled[c]
        r.num_luma_aps, r.luma_aps, r.chroma_aps = q.num_luma_aps, q.luma_aps, q.chroma_aps
        keep = []
        for c in range(3):
            if en[c] is not None:
                m = np.where(mine, en[c], 0).astype(np.uint8)
                keep.append(m)
                r.ctu_enable[c] = m.ctypes.data_as(C.POINTER(C.c_uint8))
        for c in range(2):
            r.ccalf_enabled[c] = q.ccalf_enabled[c]
            if idc[c] is not None:
                m = np.where(mine, idc[c], 0).astype(np.uint8)
                keep.append(m)
                r.ccalf_idc[c] = m.ctypes.data_as(C.POINTER(C.c_uint8))
        # CTUs of other slices may point at filter sets / alternatives this slice does not have: they are off, keep the indices legal
        if p0.ctu_filter_idx:
            f = np.ctypeslib.as_array(p0.ctu_filter_idx, shape=(n,)).copy()
            f[~mine] = 0
            keep.append(f)
            r.ctu_filter_idx = f.ctypes.data_as(C.POINTER(C.c_int16))
        for c in range(2):
            if p0.ctu_alt[c]:
                a = np.ctypeslib.as_array(p0.ctu_alt[c], shape=(n,)).copy()
                a[~mine] = 0
                keep.append(a)
                r.ctu_alt[c] = a.ctypes.data_as(C.POINTER(C.c_uint8))
        run = [p.copy() for p in src]
        alf(seq, run, r)
        for a in np.nonzero(mine)[0]:
            x0, y0 = (a % wc) * ctu, (a // wc) * ctu
            planes[0][y0:y0 + ctu, x0:x0 + ctu] = run[0][y0:y0 + ctu, x0:x0 + ctu]
            for c in range(1, len(planes)):
                planes[c][y0 >> sy:(y0 + ctu) >> sy, x0 >> sx:(x0 + ctu) >> sx] = run[c][y0 >> sy:(y0 + ctu) >> sy, x0 >> sx:(x0 + ctu) >> sx]


def filter_capture(cap, stages=("dbf", "sao", "alf")):
    """Runs the chain of the oracle on a vvc_b200.capture.Capture; returns {stage: [planes]} after each stage."""
    planes = [p.copy() for p in cap.pre]
    out = {}
    if "dbf" in stages:
        deblock(cap.seq, planes, cap.deblock_params())
        out["dbf"] = [p.copy() for p in planes]
    ctus = cap.sao_ctus()
    if "sao" in stages and ctus is not None:
        sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, cap.sao_scale[0], cap.sao_scale[1])
        sao(cap.seq, planes, ctus, cap.vb_struct())
        out["sao"] = [p.copy() for p in planes]
    ap = cap.alf_params()
    if "alf" in stages and ap is not None:
        alf(cap.seq, planes, ap)
        out["alf"] = [p.copy() for p in planes]
    out["final"] = planes
    return out
