"""ctypes binding of libvtmgpu.so (vvc_b200/csrc, CUDA for sm_100a) and the Python mirror of the reference's
operator interface for the in-loop filter chain.

Reference interface mirrored (names and call order as in DecLib::executeLoopFilters, DecoderLib/DecLib.cpp:560-623):

    LoopFilter.loopFilterPic            CommonLib/LoopFilter.cpp:145
    SampleAdaptiveOffset.SAOProcess     CommonLib/SampleAdaptiveOffset.cpp:618
    AdaptiveLoopFilter.ALFProcess       CommonLib/AdaptiveLoopFilter.cpp:393

There is no CPU path in here: if the library is missing, cannot be loaded or finds no CUDA device, every entry
point raises VtmGpuError.
"""
import ctypes as C
import os

import numpy as np

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libvtmgpu.so")


class VtmGpuError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """Loads libvtmgpu.so and declares the prototypes of include/vtmgpu.h.  Raises VtmGpuError when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("VTMGPU_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise VtmGpuError("libvtmgpu.so not found at %s -- build it with __graft_entry__.build() (there is no CPU fallback)" % p)
    try:
        lib = C.CDLL(p)
    except OSError as e:
        raise VtmGpuError("cannot load %s: %s (there is no CPU fallback)" % (p, e))
    ctx = C.c_void_p
    planes_in = [ctx, C.c_int, abi.PlanePtrs, abi.Strides]
    proto = {
        "vtmgpu_abi_version": (C.c_int, []),
        "vtmgpu_abi_sizeof": (C.c_int, [C.c_int]),
        "vtmgpu_last_error": (C.c_char_p, [ctx]),
        "vtmgpu_create": (C.c_int, [C.POINTER(abi.SeqParams), C.POINTER(ctx)]),
        "vtmgpu_destroy": (None, [ctx]),
        "vtmgpu_upload": (C.c_int, planes_in), "vtmgpu_download": (C.c_int, planes_in),
        "vtmgpu_upload_async": (C.c_int, planes_in), "vtmgpu_download_async": (C.c_int, planes_in),
        "vtmgpu_set_deblock": (C.c_int, [ctx, C.c_int, C.POINTER(abi.DeblockParams)]),
        "vtmgpu_set_deblock_async": (C.c_int, [ctx, C.c_int, C.POINTER(abi.DeblockParams)]),
        "vtmgpu_set_sao": (C.c_int, [ctx, C.c_int, C.POINTER(abi.SaoParams)]),
        "vtmgpu_set_alf": (C.c_int, [ctx, C.c_int, C.POINTER(abi.AlfParams)]),
        "vtmgpu_set_lmcs": (C.c_int, [ctx, C.c_int, C.POINTER(C.c_int16), C.c_int]),
        "vtmgpu_download_extended": (C.c_int, planes_in + [C.c_int]),
        "vtmgpu_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
        "vtmgpu_host_unregister": (C.c_int, [C.c_void_p]),
        "vtmgpu_set_alf_slices": (C.c_int, [ctx, C.c_int, C.c_int, C.POINTER(abi.AlfParams), C.POINTER(C.c_uint8)]),
        "vtmgpu_sao_reconstruct": (C.c_int, [C.POINTER(abi.SaoCtu), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "vtmgpu_deblock": (C.c_int, [ctx, C.c_int, C.c_int]), "vtmgpu_sao": (C.c_int, [ctx, C.c_int, C.c_int]),
        "vtmgpu_alf": (C.c_int, [ctx, C.c_int, C.c_int]), "vtmgpu_sao_alf": (C.c_int, [ctx, C.c_int, C.c_int]),
        "vtmgpu_deblock_sao": (C.c_int, [ctx, C.c_int, C.c_int]),
        "vtmgpu_set_deblock_sparse": (C.c_int, [ctx, C.c_int, C.POINTER(abi.DeblockSparse)]),
        "vtmgpu_set_deblock_units": (C.c_int, [ctx, C.c_int, C.c_void_p]),
        "vtmgpu_get_deblock_records": (C.c_int, [ctx, C.c_int, C.POINTER(C.c_void_p * 2), C.POINTER(C.c_void_p * 2)]),
        "vtmgpu_filter": (C.c_int, [ctx, C.c_int, C.c_int]), "vtmgpu_filter_async": (C.c_int, [ctx, C.c_int, C.c_int]),
        "vtmgpu_sync": (C.c_int, [ctx]), "vtmgpu_timer_start": (C.c_int, [ctx]),
        "vtmgpu_timer_stop": (C.c_int, [ctx, C.POINTER(C.c_float)]),
        "vtmgpu_rewind": (C.c_int, [ctx, C.c_int, C.c_int]),
        "vtmgpu_launch_count": (C.c_int64, [ctx]),
        "vtmgpu_set_profiling": (C.c_int, [ctx, C.c_int]),
        "vtmgpu_stage_ms": (C.c_int, [ctx, C.POINTER(C.c_float * 4)]),
        "vtmgpu_set_rows": (C.c_int, [ctx, C.c_int, C.c_int]),
        "vtmgpu_set_stream": (C.c_int, [ctx, C.c_void_p, C.c_int]),
        "vtmgpu_upload_rows": (C.c_int, planes_in + [C.c_int, C.c_int]), "vtmgpu_download_rows": (C.c_int, planes_in + [C.c_int, C.c_int]),
        "vtmgpu_export_rows": (C.c_int, [ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
        "vtmgpu_import_rows": (C.c_int, [ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
        "vtmgpu_export_halo": (C.c_int, [ctx, C.c_int, C.POINTER(C.c_int * 3), C.c_int, C.c_void_p]),
        "vtmgpu_import_halo": (C.c_int, [ctx, C.c_int, C.POINTER(C.c_int * 3), C.c_int, C.c_void_p]),
        "vtmgpu_hash": (C.c_int, [ctx, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int)]),
        "vtmgpu_band_export": (C.c_int, [ctx, C.POINTER(abi.BandHandle)]),
        "vtmgpu_band_connect": (C.c_int, [ctx, C.POINTER(abi.BandHandle), C.POINTER(abi.BandHandle)]),
        "vtmgpu_band_filter_async": (C.c_int, [ctx, C.c_int]),
        "vtmgpu_band_disconnect": (C.c_int, [ctx]),
        "vtmgpu_batch_create": (C.c_int, [C.POINTER(abi.SeqParams), C.c_int, C.POINTER(ctx)]),
        "vtmgpu_batch_destroy": (None, [ctx]),
        "vtmgpu_batch_filter": (C.c_int, [ctx, C.POINTER(abi.HostPicture), C.c_int]),
        "vtmgpu_batch_last_error": (C.c_char_p, [ctx]),
        "vtmgpu_batch_launch_count": (C.c_int64, [ctx]),
    }
    for name, (res, args) in proto.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise VtmGpuError("%s does not export %s" % (p, name))
        fn.restype, fn.argtypes = res, args
    if lib.vtmgpu_abi_version() != abi.ABI_VERSION:
        raise VtmGpuError("libvtmgpu ABI version %d, binding expects %d" % (lib.vtmgpu_abi_version(), abi.ABI_VERSION))
    if path is None:
        _lib = lib
    return lib


def sao_reconstruct(ctus, width_in_ctus, ncomp, scale_luma=0, scale_chroma=0):
    """Host-only helper of the library: xReconstructBlkSAOParams (SampleAdaptiveOffset.cpp:266) in place.
    Returns the 3-bit mask of components with SAO on; raises on malformed parameters."""
    rc = load_library().vtmgpu_sao_reconstruct(ctus, len(ctus), width_in_ctus, ncomp, scale_luma, scale_chroma)
    if rc < 0:
        raise VtmGpuError("vtmgpu_sao_reconstruct: invalid SAO parameters (rc=%d)" % rc)
    return rc


def sparse_records(dbf_luma, dbf_chroma, pin=False, ladf=None):
    """List form (abi.DeblockSparse) of dense record arrays: what a producer that appends while walking the CUs emits.  The four
    lists are laid out in ONE buffer (array order, each on a 16-byte boundary) so that the library uploads them with one copy.
    pin=True places that buffer in page-locked memory (torch) so that the upload is asynchronous."""
    import numpy as np
    p = abi.DeblockSparse()
    lists = []
    for d in range(2):
        dense = np.ascontiguousarray(dbf_luma[d]).reshape(-1).view(np.uint32)
        idx = np.flatnonzero(dense)
        ent = np.zeros(len(idx), dtype=abi.LUMA_ENTRY_DTYPE)
        ent["index"], ent["rec"] = idx, dense[idx]
        lists.append(ent)
    for d in range(2):
        if dbf_chroma is not None and dbf_chroma[d].size:
            dense = np.ascontiguousarray(dbf_chroma[d]).reshape(-1).view(np.uint64)
            idx = np.flatnonzero(dense)
            ent = np.zeros(len(idx), dtype=abi.CHROMA_ENTRY_DTYPE)
            ent["index"], ent["rec"] = idx, dense[idx]
        else:
            ent = np.zeros(0, dtype=abi.CHROMA_ENTRY_DTYPE)
        lists.append(ent)
    offs, total = [], 0
    for ent in lists:
        offs.append(total)
        total = (total + ent.nbytes + 15) & ~15
    if pin:
        import torch
        holder = torch.empty(max(total, 16), dtype=torch.uint8).pin_memory()
        buf, base = holder.numpy(), holder.data_ptr()
    else:
        holder = np.zeros(max(total, 16) + 16, dtype=np.uint8)
        skew = (-holder.ctypes.data) & 15
        buf, base = holder[skew:], holder.ctypes.data + skew
    for ent, off in zip(lists, offs):
        buf[off:off + ent.nbytes] = ent.view(np.uint8).reshape(-1)
    for d in range(2):
        p.luma_count[d] = len(lists[d])
        p.chroma_count[d] = len(lists[2 + d])
        if len(lists[d]):
            p.luma[d] = C.cast(base + offs[d], C.POINTER(abi.DbfLumaEntry))
        if len(lists[2 + d]):
            p.chroma[d] = C.cast(base + offs[2 + d], C.POINTER(abi.DbfChromaEntry))
    keep = [holder]
    if ladf is not None:
        keep.append(ladf)
        p.ladf = C.pointer(ladf)
    p._keep = keep
    p.nbytes = total
    return p


def _plane_args(planes):
    ptrs, strides = abi.PlanePtrs(), abi.Strides()
    for c, p in enumerate(planes):
        if p.dtype != np.int16 or p.ndim != 2 or p.strides[1] != 2:
            raise ValueError("planes must be 2-D int16 arrays with contiguous rows")
        ptrs[c] = C.cast(p.ctypes.data, C.POINTER(C.c_int16))
        strides[c] = p.strides[0] // 2
    return ptrs, strides


def host_picture(planes_in, planes_out, deblock=None, sao_ctus=None, alf=None, vb=None):
    """abi.HostPicture for vtmgpu_batch_filter from numpy planes (C-contiguous int16; page-locked memory makes the copies
    asynchronous), an abi.DeblockSparse, a reconstructed (abi.SaoCtu * n) array and an abi.AlfParams.  The returned object keeps
    everything it points to alive."""
    hp = abi.HostPicture()
    keep = [planes_in, planes_out, deblock, sao_ctus, alf, vb]
    for k, (a, b) in enumerate(zip(planes_in, planes_out)):
        assert a.dtype == np.int16 and b.dtype == np.int16 and a.flags["C_CONTIGUOUS"] and b.flags["C_CONTIGUOUS"]
        hp.inp[k] = a.ctypes.data_as(C.POINTER(C.c_int16))
        hp.out[k] = b.ctypes.data_as(C.POINTER(C.c_int16))
        hp.in_stride[k] = a.strides[0] // 2
        hp.out_stride[k] = b.strides[0] // 2
    if deblock is not None:
        hp.deblock = C.pointer(deblock)
    if sao_ctus is not None:
        sp = abi.SaoParams(C.cast(sao_ctus, C.POINTER(abi.SaoCtu)), len(sao_ctus), C.pointer(vb) if vb is not None else None)
        keep.append(sp)
        hp.sao = C.pointer(sp)
    if alf is not None:
        hp.alf = C.pointer(alf)
    hp._keep = keep
    return hp


class Batch:
    """vtmgpu_batch: `lanes` single-picture contexts; filter() takes a run of host pictures through the whole boundary in ONE
    C call (copies of one picture overlap the kernels of another; one issuing thread)."""

    def __init__(self, seq, lanes=4, device=0):
        self.lib = load_library()
        sp = abi.SeqParams(seq["width"], seq["height"], seq["chroma_format"], seq["bit_depth_luma"], seq["bit_depth_chroma"], seq["ctu_size"], 1, device)
        self.h = C.c_void_p()
        if self.lib.vtmgpu_batch_create(C.byref(sp), lanes, C.byref(self.h)):
            raise VtmGpuError(self.lib.vtmgpu_last_error(None).decode())

    def filter(self, pictures):
        """pictures: list of abi.HostPicture (see host_picture) or a ready (abi.HostPicture * n) array."""
        arr = pictures if isinstance(pictures, C.Array) else (abi.HostPicture * len(pictures))(*pictures)
        if self.lib.vtmgpu_batch_filter(self.h, arr, len(arr)):
            raise VtmGpuError(self.lib.vtmgpu_batch_last_error(self.h).decode())

    def launch_count(self):
        return int(self.lib.vtmgpu_batch_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.vtmgpu_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One libvtmgpu context: device planes + side info for `capacity` picture slots on one GPU."""

    def __init__(self, seq, capacity=1, device=0):
        self.lib = load_library()
        self.seq = dict(seq)
        sp = abi.SeqParams(seq["width"], seq["height"], seq["chroma_format"], seq["bit_depth_luma"], seq["bit_depth_chroma"],
                           seq["ctu_size"], capacity, device)
        self.h = C.c_void_p()
        if self.lib.vtmgpu_create(C.byref(sp), C.byref(self.h)):
            raise VtmGpuError(self.lib.vtmgpu_last_error(None).decode())
        self.capacity = capacity
        self.device = device
        self.ncomp = 1 if seq["chroma_format"] == 0 else 3
        sx, sy = abi.chroma_shifts(seq["chroma_format"])
        w, h = seq["width"], seq["height"]
        self.shapes = [(h, w)] + ([(h >> sy, w >> sx)] * 2 if self.ncomp == 3 else [])

    def close(self):
        if self.h:
            self.lib.vtmgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise VtmGpuError("%s: %s" % (what, self.lib.vtmgpu_last_error(self.h).decode()))

    # ---- planes ----------------------------------------------------------------------------------------
    def upload(self, slot, planes, sync=True):
        ptrs, strides = _plane_args(planes)
        self._ck((self.lib.vtmgpu_upload if sync else self.lib.vtmgpu_upload_async)(self.h, slot, ptrs, strides), "upload")

    def download(self, slot, out=None, sync=True):
        if out is None:
            out = [np.empty(s, dtype=np.int16) for s in self.shapes]
        ptrs, strides = _plane_args(out)
        self._ck((self.lib.vtmgpu_download if sync else self.lib.vtmgpu_download_async)(self.h, slot, ptrs, strides), "download")
        return out

    # ---- band mode (one picture over several contexts) -----------------------------------------------------
    def set_rows(self, y_begin, y_end):
        self._ck(self.lib.vtmgpu_set_rows(self.h, y_begin, y_end), "set_rows")

    def set_stream(self, cuda_stream, async_stages=True):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or None for the context's own stream."""
        self._ck(self.lib.vtmgpu_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None, int(async_stages)), "set_stream")

    def upload_rows(self, slot, planes, y_begin, y_end):
        """planes = the FULL host picture; only luma rows [y_begin, y_end) (and the collocated chroma rows) are copied."""
        ptrs, strides = _plane_args(planes)
        self._ck(self.lib.vtmgpu_upload_rows(self.h, slot, ptrs, strides, y_begin, y_end), "upload_rows")

    def download_rows(self, slot, planes, y_begin, y_end):
        ptrs, strides = _plane_args(planes)
        self._ck(self.lib.vtmgpu_download_rows(self.h, slot, ptrs, strides, y_begin, y_end), "download_rows")

    def export_rows(self, slot, comp, y0, nrows, dev_ptr):
        self._ck(self.lib.vtmgpu_export_rows(self.h, slot, comp, y0, nrows, C.c_void_p(dev_ptr)), "export_rows")

    def import_rows(self, slot, comp, y0, nrows, dev_ptr):
        self._ck(self.lib.vtmgpu_import_rows(self.h, slot, comp, y0, nrows, C.c_void_p(dev_ptr)), "import_rows")

    def export_halo(self, slot, rows, nrows, dev_ptr):
        """rows = first row per component; all planes packed into one device buffer."""
        y = (C.c_int * 3)(*(list(rows) + [0, 0])[:3])
        self._ck(self.lib.vtmgpu_export_halo(self.h, slot, C.byref(y), nrows, C.c_void_p(dev_ptr)), "export_halo")

    def import_halo(self, slot, rows, nrows, dev_ptr):
        y = (C.c_int * 3)(*(list(rows) + [0, 0])[:3])
        self._ck(self.lib.vtmgpu_import_halo(self.h, slot, C.byref(y), nrows, C.c_void_p(dev_ptr)), "import_halo")

    # ---- decoded-picture hash on the device -----------------------------------------------------------------
    def hash(self, first=0, count=1, kind=abi.HASH_MD5):
        """Digests of the current planes of slots [first, first + count) computed on the device: list (per slot) of bytes holding
        the components' digests one after the other, as in the decoded picture hash SEI (kind: abi.HASH_MD5 / HASH_CRC / HASH_CHECKSUM)."""
        buf = (C.c_uint8 * (abi.HASH_SLOT_BYTES * count))()
        n = C.c_int(0)
        self._ck(self.lib.vtmgpu_hash(self.h, first, count, kind, buf, C.byref(n)), "hash")
        raw = bytes(buf)
        return [raw[s * abi.HASH_SLOT_BYTES:s * abi.HASH_SLOT_BYTES + n.value * self.ncomp] for s in range(count)]

    # ---- band mode over peer memory ---------------------------------------------------------------------
    def band_export(self):
        """bytes of this rank's vtmgpu_band_handle (to be handed to the neighbouring ranks by any host channel)."""
        h = abi.BandHandle()
        self._ck(self.lib.vtmgpu_band_export(self.h, C.byref(h)), "band_export")
        return bytes(h.bytes)

    def band_connect(self, above=None, below=None):
        """above / below: handle bytes of the ranks that own the neighbouring bands, or None at the picture border."""
        hs = [abi.BandHandle.from_buffer_copy(b) if b is not None else None for b in (above, below)]
        self._ck(self.lib.vtmgpu_band_connect(self.h, C.byref(hs[0]) if hs[0] is not None else None, C.byref(hs[1]) if hs[1] is not None else None), "band_connect")

    def band_filter(self, slot=0):
        """Enqueues the whole chain for this rank's band with the halo rows stored straight into the neighbours' planes."""
        self._ck(self.lib.vtmgpu_band_filter_async(self.h, slot), "band_filter")

    def band_disconnect(self):
        self._ck(self.lib.vtmgpu_band_disconnect(self.h), "band_disconnect")

    # ---- side information --------------------------------------------------------------------------------
    def set_deblock(self, slot, params, sync=True):
        """sync=False: no staging copy, the record arrays (page-locked) are read asynchronously until the next sync()."""
        fn = self.lib.vtmgpu_set_deblock if sync else self.lib.vtmgpu_set_deblock_async
        self._ck(fn(self.h, slot, C.byref(params) if params is not None else None), "set_deblock")

    def set_deblock_sparse(self, slot, params):
        """params: abi.DeblockSparse (see sparse_records); the lists are read asynchronously until the next sync()."""
        self._ck(self.lib.vtmgpu_set_deblock_sparse(self.h, slot, C.byref(params) if params is not None else None), "set_deblock_sparse")

    def set_sao(self, slot, ctus, vb=None):
        """ctus: reconstructed (abi.SaoCtu * n) array, or None to switch the stage off; vb: abi.VirtualBoundaries or None."""
        if ctus is None:
            self._ck(self.lib.vtmgpu_set_sao(self.h, slot, None), "set_sao")
            return
        p = abi.SaoParams(C.cast(ctus, C.POINTER(abi.SaoCtu)), len(ctus), C.pointer(vb) if vb is not None else None)
        self._ck(self.lib.vtmgpu_set_sao(self.h, slot, C.byref(p)), "set_sao")

    def set_lmcs(self, slot, inv_lut):
        """inv_lut: int16 numpy array of 1 << bit_depth_luma entries (Reshape::getInvLUT()), or None = off.  The luma plane uploaded to
        the slot is then the reshaped-domain reconstruction; the deblocking / SAO pass maps it while it loads its tiles."""
        if inv_lut is None:
            self._ck(self.lib.vtmgpu_set_lmcs(self.h, slot, None, 0), "set_lmcs")
            return
        lut = np.ascontiguousarray(inv_lut, dtype=np.int16)
        self._ck(self.lib.vtmgpu_set_lmcs(self.h, slot, lut.ctypes.data_as(C.POINTER(C.c_int16)), len(lut)), "set_lmcs")

    def download_extended(self, slot, margin_luma):
        """Returns the planes WITH their margins (Picture::extendPicBorder): arrays of (h + 2 ym, w + 2 xm)."""
        sx, sy = abi.chroma_shifts(self.seq["chroma_format"])
        out, ptrs, strides = [], abi.PlanePtrs(), abi.Strides()
        for c in range(self.ncomp):
            xm, ym = (margin_luma >> sx, margin_luma >> sy) if c else (margin_luma, margin_luma)
            w, h = (self.seq["width"] >> sx, self.seq["height"] >> sy) if c else (self.seq["width"], self.seq["height"])
            a = np.full((h + 2 * ym, w + 2 * xm), -1, dtype=np.int16)
            out.append(a)
            ptrs[c] = C.cast(a.ctypes.data + 2 * (ym * a.shape[1] + xm), C.POINTER(C.c_int16))
            strides[c] = a.shape[1]
        self._ck(self.lib.vtmgpu_download_extended(self.h, slot, ptrs, strides, margin_luma), "download_extended")
        return out

    def set_alf(self, slot, params):
        self._ck(self.lib.vtmgpu_set_alf(self.h, slot, C.byref(params) if params is not None else None), "set_alf")

    def set_alf_slices(self, slot, slices, ctu_slice):
        """slices: list of abi.AlfParams, one per slice of the picture (per-picture arrays are read from slices[0]);
        ctu_slice: uint8 numpy array, slice index per CTU (ALFProcess reloads the APS data at every slice change)."""
        arr = (abi.AlfParams * len(slices))()
        for i, q in enumerate(slices):
            C.memmove(C.byref(arr[i]), C.byref(q), C.sizeof(abi.AlfParams))
        self._ck(self.lib.vtmgpu_set_alf_slices(self.h, slot, len(slices), arr, ctu_slice.ctypes.data_as(C.POINTER(C.c_uint8))), "set_alf_slices")

    def set_capture(self, slot, cap, upload=True, sync=True):
        """Loads one captured picture (planes + all side information) into a slot."""
        if upload:
            self.upload(slot, cap.pre, sync=sync)
        self.set_deblock(slot, cap.deblock_params())
        ctus = cap.sao_ctus()
        if ctus is not None:
            sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, cap.sao_scale[0], cap.sao_scale[1])
        self.set_sao(slot, ctus, cap.vb_struct())
        self.set_alf(slot, cap.alf_params())

    # ---- stages ------------------------------------------------------------------------------------------
    def deblock(self, first=0, count=1):
        self._ck(self.lib.vtmgpu_deblock(self.h, first, count), "deblock")

    def sao(self, first=0, count=1):
        self._ck(self.lib.vtmgpu_sao(self.h, first, count), "sao")

    def alf(self, first=0, count=1):
        self._ck(self.lib.vtmgpu_alf(self.h, first, count), "alf")

    def deblock_sao(self, first=0, count=1):
        self._ck(self.lib.vtmgpu_deblock_sao(self.h, first, count), "deblock_sao")

    def sao_alf(self, first=0, count=1):
        self._ck(self.lib.vtmgpu_sao_alf(self.h, first, count), "sao_alf")

    def filter(self, first=0, count=1, sync=True):
        self._ck((self.lib.vtmgpu_filter if sync else self.lib.vtmgpu_filter_async)(self.h, first, count), "filter")

    def sync(self):
        self._ck(self.lib.vtmgpu_sync(self.h), "sync")

    def rewind(self, first=0, count=1):
        self._ck(self.lib.vtmgpu_rewind(self.h, first, count), "rewind")

    def timer_start(self):
        self._ck(self.lib.vtmgpu_timer_start(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.lib.vtmgpu_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def launch_count(self):
        return int(self.lib.vtmgpu_launch_count(self.h))

    def set_profiling(self, on):
        self._ck(self.lib.vtmgpu_set_profiling(self.h, int(on)), "set_profiling")

    def stage_ms(self):
        ms = (C.c_float * 4)()
        self._ck(self.lib.vtmgpu_stage_ms(self.h, C.byref(ms)), "stage_ms")
        return [float(v) for v in ms]


# ---------------------------------------------------------------------------------------------------------------
# mirror of the reference's three filter classes (same names / call order as DecLib::executeLoopFilters)
# ---------------------------------------------------------------------------------------------------------------
class Picture:
    """What DecLib hands to the filters, flattened: reco planes (modified in place) + the side information."""

    def __init__(self, cap):
        self.cap = cap
        self.seq = cap.seq
        self.reco = [p.copy() for p in cap.pre]


class LoopFilter:
    """LoopFilter (CommonLib/LoopFilter.h:107-129)."""

    def __init__(self):
        self.ctx = None

    def create(self, seq, device=0):
        if self.ctx is None or self.ctx.seq != dict(seq):
            self.ctx = Context(seq, capacity=1, device=device)

    def destroy(self):
        if self.ctx:
            self.ctx.close()
            self.ctx = None

    def loopFilterPic(self, pic):
        self.ctx.upload(0, pic.reco)
        self.ctx.set_deblock(0, pic.cap.deblock_params())
        self.ctx.deblock(0, 1)
        self.ctx.download(0, pic.reco)


class SampleAdaptiveOffset:
    """SampleAdaptiveOffset (CommonLib/SampleAdaptiveOffset.h:63-73); shares the LoopFilter's device context."""

    def __init__(self, loop_filter):
        self.lf = loop_filter

    def SAOProcess(self, pic, sao_ctus):
        if sao_ctus is None:
            raise VtmGpuError("No parameters present")           # CHECK at SampleAdaptiveOffset.cpp:621
        sao_reconstruct(sao_ctus, pic.cap.width_in_ctus, pic.cap.ncomp, pic.cap.sao_scale[0], pic.cap.sao_scale[1])
        self.lf.ctx.set_sao(0, sao_ctus, pic.cap.vb_struct())
        self.lf.ctx.sao(0, 1)
        self.lf.ctx.download(0, pic.reco)


class AdaptiveLoopFilter:
    """AdaptiveLoopFilter (CommonLib/AdaptiveLoopFilter.h:83-120)."""

    def __init__(self, loop_filter):
        self.lf = loop_filter

    def ALFProcess(self, pic):
        self.lf.ctx.set_alf(0, pic.cap.alf_params())
        self.lf.ctx.alf(0, 1)
        self.lf.ctx.download(0, pic.reco)


def execute_loop_filters(cap, device=0, fused=True, ctx=None):
    """DecLib::executeLoopFilters for one captured picture.  Returns {stage: planes} (fused: only 'final')."""
    own = ctx is None
    if own:
        ctx = Context(cap.seq, capacity=1, device=device)
    try:
        out = {}
        if fused:
            ctx.set_capture(0, cap)
            ctx.filter(0, 1)
            out["final"] = ctx.download(0)
            return out
        pic = Picture(cap)
        lf = LoopFilter()
        lf.ctx = ctx
        ctx.set_sao(0, None)
        ctx.set_alf(0, None)
        lf.loopFilterPic(pic)
        out["dbf"] = [p.copy() for p in pic.reco]
        ctus = cap.sao_ctus()
        if ctus is not None:
            SampleAdaptiveOffset(lf).SAOProcess(pic, ctus)
            out["sao"] = [p.copy() for p in pic.reco]
        if cap.alf is not None:
            AdaptiveLoopFilter(lf).ALFProcess(pic)
            out["alf"] = [p.copy() for p in pic.reco]
        out["final"] = pic.reco
        return out
    finally:
        if own:
            ctx.close()
