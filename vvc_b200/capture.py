"""Reader for per-picture captures of the drop-in boundary (vvc_b200/shim/capture_format.h, "VTMGCAP1")
and their compressed .npz form used for the committed golden fixtures (tests/golden/).

A Capture holds everything that crosses the boundary for one picture: the pre-filter planes, the deblocking
segment records, the SAO CTU parameters as parsed, the ALF slice/APS/CTU data -- and, when the capture was taken
with a staged backend, the planes after each stage (deblocked / SAO / ALF = final).
"""
import ctypes as C
import io
import os
import struct

import numpy as np

from . import abi

STAGES = ("dbf", "sao", "alf")


def _plane_shapes(seq):
    sx, sy = abi.chroma_shifts(seq["chroma_format"])
    w, h = seq["width"], seq["height"]
    if seq["chroma_format"] == 0:
        return [(h, w)]
    return [(h, w), (h >> sy, w >> sx), (h >> sy, w >> sx)]


class Capture:
    def __init__(self, sections):
        s = sections
        hdr = np.frombuffer(s["seq"], dtype=np.int32)
        self.seq = dict(width=int(hdr[0]), height=int(hdr[1]), chroma_format=int(hdr[2]), bit_depth_luma=int(hdr[3]),
                        bit_depth_chroma=int(hdr[4]), ctu_size=int(hdr[5]), poc=int(hdr[6]), stages_mask=int(hdr[7]))
        shapes = _plane_shapes(self.seq)
        self.ncomp = len(shapes)

        def planes(prefix):
            if prefix + "_0" not in s:
                return None
            return [np.frombuffer(s["%s_%d" % (prefix, c)], dtype=np.int16).reshape(shapes[c]).copy() for c in range(self.ncomp)]

        self.pre = planes("pre")
        self.stage = {k: planes(k) for k in STAGES}
        self.dbf_luma = [np.frombuffer(s["dbfrec_l%d" % d], dtype=np.uint32).copy() for d in range(2)]
        self.dbf_chroma = [np.frombuffer(s["dbfrec_c%d" % d], dtype=np.uint64).copy() for d in range(2)]
        self.ladf = np.frombuffer(s["dbf_ladf"], dtype=np.int32).copy() if s.get("dbf_ladf") else None    # vtmgpu_ladf as 11 int32
        self.vb = np.frombuffer(s["vb"], dtype=np.int32).copy() if s.get("vb") else None          # vtmgpu_virtual_boundaries as 8 int32
        self.sao_raw = bytes(s["sao_raw"]) if "sao_raw" in s else None
        self.sao_scale = [int(v) for v in np.frombuffer(s["sao_scale"], dtype=np.int32)] if "sao_scale" in s else [0, 0]
        self.alf = None
        if "alf_hdr" in s:
            h = np.frombuffer(s["alf_hdr"], dtype=np.int32)
            self.alf = dict(enabled=[int(v) for v in h[0:3]], num_luma_aps=int(h[3]), has_chroma_aps=int(h[4]),
                            ccalf_enabled=[int(h[5]), int(h[6])], num_ctus=int(h[7]),
                            luma_aps=bytes(s["alf_luma_aps"]), chroma_aps=bytes(s["alf_chroma_aps"]),
                            ctu_enable=[np.frombuffer(s["alf_en%d" % c], dtype=np.uint8).copy() for c in range(3)],
                            filter_idx=np.frombuffer(s["alf_fidx"], dtype=np.int16).copy(),
                            ctu_alt=[np.frombuffer(s["alf_alt%d" % c], dtype=np.uint8).copy() for c in range(2)],
                            cc_coeff=np.frombuffer(s["alf_cccoef"], dtype=np.int16).reshape(2, 4, 8).copy(),
                            cc_idc=[np.frombuffer(s["alf_ccidc%d" % c], dtype=np.uint8).copy() for c in range(2)],
                            clip=np.frombuffer(s["alf_clip"], dtype=np.uint8).copy() if s.get("alf_clip") else None)
        self._sections = s

    # ---- geometry -------------------------------------------------------------------------------------
    @property
    def width(self):
        return self.seq["width"]

    @property
    def height(self):
        return self.seq["height"]

    @property
    def num_ctus(self):
        c = self.seq["ctu_size"]
        return ((self.width + c - 1) // c) * ((self.height + c - 1) // c)

    @property
    def width_in_ctus(self):
        c = self.seq["ctu_size"]
        return (self.width + c - 1) // c

    def luma_pixels(self):
        return self.width * self.height

    # ---- ctypes views (the returned objects keep the backing arrays alive through ._keep) ---------------
    def deblock_params(self):
        p = abi.DeblockParams()
        keep = []
        for d in range(2):
            a = np.ascontiguousarray(self.dbf_luma[d])
            keep.append(a)
            p.luma[d] = a.ctypes.data_as(C.POINTER(C.c_uint32))
            if self.dbf_chroma[d].size:
                b = np.ascontiguousarray(self.dbf_chroma[d])
                keep.append(b)
                p.chroma[d] = b.ctypes.data_as(C.POINTER(C.c_uint64))
        lad = self.ladf_struct()
        if lad is not None:
            keep.append(lad)
            p.ladf = C.pointer(lad)
        p._keep = keep
        return p

    def vb_struct(self):
        """abi.VirtualBoundaries signalled for the picture, or None."""
        return abi.VirtualBoundaries.from_buffer_copy(self.vb.tobytes()) if self.vb is not None else None

    def ladf_struct(self):
        """abi.Ladf of the sequence, or None when LADF is off (then the luma records carry tc / beta)."""
        return abi.Ladf.from_buffer_copy(self.ladf.tobytes()) if self.ladf is not None else None

    def sao_ctus(self):
        """Fresh, writable array of SaoCtu as parsed (NOT yet reconstructed); None when the stream has no SAO stage."""
        if self.sao_raw is None:
            return None
        n = len(self.sao_raw) // C.sizeof(abi.SaoCtu)
        arr = (abi.SaoCtu * n).from_buffer_copy(self.sao_raw)
        return arr

    def alf_params(self):
        if self.alf is None:
            return None
        a = self.alf
        p = abi.AlfParams()
        keep = []
        for c in range(3):
            p.enabled[c] = a["enabled"][c]
        p.num_luma_aps = a["num_luma_aps"]
        if a["num_luma_aps"]:
            la = (abi.AlfLumaAps * a["num_luma_aps"]).from_buffer_copy(a["luma_aps"])
            keep.append(la)
            p.luma_aps = C.cast(la, C.POINTER(abi.AlfLumaAps))
        if a["has_chroma_aps"]:
            ca = abi.AlfChromaAps.from_buffer_copy(a["chroma_aps"])
            keep.append(ca)
            p.chroma_aps = C.pointer(ca)
        for c in range(3):
            p.ctu_enable[c] = a["ctu_enable"][c].ctypes.data_as(C.POINTER(C.c_uint8))
        p.ctu_filter_idx = a["filter_idx"].ctypes.data_as(C.POINTER(C.c_int16))
        for c in range(2):
            p.ctu_alt[c] = a["ctu_alt"][c].ctypes.data_as(C.POINTER(C.c_uint8))
            p.ccalf_enabled[c] = a["ccalf_enabled"][c]
            p.ccalf_idc[c] = a["cc_idc"][c].ctypes.data_as(C.POINTER(C.c_uint8))
            for f in range(4):
                for k in range(8):
                    p.ccalf_coeff[c][f][k] = int(a["cc_coeff"][c, f, k])
        p.num_ctus = a["num_ctus"]
        if a.get("clip") is not None:
            p.ctu_clip = a["clip"].ctypes.data_as(C.POINTER(C.c_uint8))
        vb = self.vb_struct()
        if vb is not None:
            keep.append(vb)
            p.vb = C.pointer(vb)
        p._keep = keep + [a]
        return p

    # ---- activity statistics (what work a picture actually contains) -------------------
    def activity(self):
        out = {}
        for d in range(2):
            out["dbf_luma_segments_dir%d" % d] = int(np.count_nonzero(self.dbf_luma[d] & 0x7FF))
            out["dbf_chroma_segments_dir%d" % d] = int(np.count_nonzero(self.dbf_chroma[d] & 0x3FFFFF))
        out["dbf_luma_active_frac"] = float(sum(out["dbf_luma_segments_dir%d" % d] for d in range(2)) / max(1, 2 * self.dbf_luma[0].size))
        if self.sao_raw is not None:
            ctus = self.sao_ctus()
            for c in range(self.ncomp):
                out["sao_on_frac_c%d" % c] = float(np.mean([ctus[i].comp[c].mode != 0 for i in range(len(ctus))]))
        if self.alf is not None:
            for c in range(3):
                out["alf_on_frac_c%d" % c] = float(np.mean(self.alf["ctu_enable"][c] != 0))
            for c in range(2):
                out["ccalf_on_frac_c%d" % (c + 1)] = float(np.mean(self.alf["cc_idc"][c] != 0)) if self.alf["ccalf_enabled"][c] else 0.0
        return out

    # ---- compressed fixture form ----------------------------------------------------------------------
    def save_npz(self, path, stages=True):
        """Stage planes are stored as int16 differences to the previous stage (they compress ~10x better)."""
        d = {}
        for k, v in self._sections.items():
            if k[:4] in ("pre_", "dbf_", "sao_", "alf_") and k[4:].isdigit():
                continue
            d["raw__" + k] = np.frombuffer(v, dtype=np.uint8)
        prev = self.pre
        for c in range(self.ncomp):
            d["pre_%d" % c] = self.pre[c]
        if stages:
            for st in STAGES:
                if self.stage[st] is None:
                    continue
                for c in range(self.ncomp):
                    d["delta_%s_%d" % (st, c)] = (self.stage[st][c] - prev[c]).astype(np.int16)
                prev = self.stage[st]
        np.savez_compressed(path, **d)


def _read_cap(path):
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"VTMGCAP1":
        raise ValueError("%s: not a VTMGCAP1 capture" % path)
    (n,) = struct.unpack_from("<I", data, 8)
    off, secs = 12, {}
    for _ in range(n):
        name = data[off:off + 24].split(b"\0", 1)[0].decode()
        (nb,) = struct.unpack_from("<Q", data, off + 24)
        off += 32
        secs[name] = data[off:off + nb]
        off += nb + (8 - nb % 8) % 8
    return secs


def _read_npz(path):
    z = np.load(path)
    secs = {}
    for k in z.files:
        if k.startswith("raw__"):
            secs[k[5:]] = z[k].tobytes()
    ncomp = sum(1 for k in z.files if k.startswith("pre_"))
    prev = [z["pre_%d" % c].astype(np.int16) for c in range(ncomp)]
    for c in range(ncomp):
        secs["pre_%d" % c] = prev[c].tobytes()
    for st in STAGES:
        if "delta_%s_0" % st not in z.files:
            continue
        cur = [(prev[c] + z["delta_%s_%d" % (st, c)]).astype(np.int16) for c in range(ncomp)]
        for c in range(ncomp):
            secs["%s_%d" % (st, c)] = cur[c].tobytes()
        prev = cur
    return secs


def load(path):
    """Loads a .cap (raw VTMGCAP1) or .npz (compressed fixture) capture."""
    return Capture(_read_npz(path) if path.endswith(".npz") else _read_cap(path))


def load_dir(path, limit=None):
    names = sorted(n for n in os.listdir(path) if n.endswith(".cap") or n.endswith(".npz"))
    if limit:
        names = names[:limit]
    return [load(os.path.join(path, n)) for n in names]
