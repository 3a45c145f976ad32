"""ctypes mirror of include/vtmgpu.h (the C ABI of libvtmgpu) -- structures and constants only.

Keep in sync with the header; tests/test_abi.py checks sizes/offsets against the compiled library
(vtmgpu_abi_sizeof) and that every declared symbol is exported.
"""
import ctypes as C

ABI_VERSION = 3

ALF_CLASSES, ALF_LUMA_COEFF, ALF_CHROMA_COEFF = 25, 13, 7
ALF_MAX_APS, ALF_MAX_ALTS, ALF_FIXED_SETS = 8, 8, 16
CCALF_MAX_FILTERS, CCALF_COEFF = 4, 8

SAO_MODE_OFF, SAO_MODE_NEW, SAO_MODE_MERGE = 0, 1, 2
SAO_EO_0, SAO_EO_90, SAO_EO_135, SAO_EO_45, SAO_BO = 0, 1, 2, 3, 4
SAO_MERGE_LEFT, SAO_MERGE_ABOVE = 0, 1

AVAIL_LEFT, AVAIL_RIGHT, AVAIL_ABOVE, AVAIL_BELOW = 0x01, 0x02, 0x04, 0x08
AVAIL_ABOVE_LEFT, AVAIL_ABOVE_RIGHT, AVAIL_BELOW_LEFT, AVAIL_BELOW_RIGHT = 0x10, 0x20, 0x40, 0x80
AVAIL_ALL = 0xFF

# deblocking record packing
DBF_L_BETA_SHIFT, DBF_L_LENP_SHIFT, DBF_L_LENQ_SHIFT = 11, 22, 25
DBF_L_PNOFILT, DBF_L_QNOFILT, DBF_L_CTUROW = 1 << 28, 1 << 29, 1 << 30
DBF_C_TCCR_SHIFT, DBF_C_BETACB_SHIFT, DBF_C_BETACR_SHIFT = 11, 22, 33
DBF_C_LARGE, DBF_C_CTB, DBF_C_PNOFILT, DBF_C_QNOFILT = 1 << 44, 1 << 45, 1 << 46, 1 << 47


class SeqParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("chroma_format", C.c_int32),
                ("bit_depth_luma", C.c_int32), ("bit_depth_chroma", C.c_int32), ("ctu_size", C.c_int32),
                ("capacity", C.c_int32), ("device", C.c_int32)]


class Ladf(C.Structure):
    _fields_ = [("num_intervals", C.c_int32), ("qp_offset", C.c_int32 * 5), ("lower_bound", C.c_int32 * 5)]


DBF_LADF_BIAS = 128


class DeblockParams(C.Structure):
    _fields_ = [("luma", C.POINTER(C.c_uint32) * 2), ("chroma", C.POINTER(C.c_uint64) * 2), ("ladf", C.POINTER(Ladf))]


class DbfLumaEntry(C.Structure):
    _fields_ = [("index", C.c_uint32), ("rec", C.c_uint32)]


class DbfChromaEntry(C.Structure):
    _fields_ = [("rec", C.c_uint64), ("index", C.c_uint32), ("reserved", C.c_uint32)]


class DeblockSparse(C.Structure):
    _fields_ = [("luma", C.POINTER(DbfLumaEntry) * 2), ("chroma", C.POINTER(DbfChromaEntry) * 2),
                ("luma_count", C.c_uint32 * 2), ("chroma_count", C.c_uint32 * 2), ("ladf", C.POINTER(Ladf))]


LUMA_ENTRY_DTYPE = [("index", "<u4"), ("rec", "<u4")]
CHROMA_ENTRY_DTYPE = [("rec", "<u8"), ("index", "<u4"), ("reserved", "<u4")]


class SaoOffset(C.Structure):
    _fields_ = [("mode", C.c_int8), ("type", C.c_int8), ("aux", C.c_int8), ("reserved", C.c_int8),
                ("offset", C.c_int16 * 32)]


class SaoCtu(C.Structure):
    _fields_ = [("comp", SaoOffset * 3), ("avail", C.c_uint8), ("merge_left_ok", C.c_uint8),
                ("merge_above_ok", C.c_uint8), ("reserved", C.c_uint8)]


class VirtualBoundaries(C.Structure):
    _fields_ = [("num_ver", C.c_int32), ("num_hor", C.c_int32), ("pos_x", C.c_int32 * 3), ("pos_y", C.c_int32 * 3)]


class SaoParams(C.Structure):
    _fields_ = [("ctu", C.POINTER(SaoCtu)), ("num_ctus", C.c_int32), ("vb", C.POINTER(VirtualBoundaries))]


class AlfLumaAps(C.Structure):
    _fields_ = [("num_filters", C.c_int32), ("nonlinear", C.c_int32), ("delta_idx", C.c_int16 * ALF_CLASSES),
                ("coeff", (C.c_int16 * ALF_LUMA_COEFF) * ALF_CLASSES),
                ("clip_idx", (C.c_int16 * ALF_LUMA_COEFF) * ALF_CLASSES)]


class AlfChromaAps(C.Structure):
    _fields_ = [("num_alts", C.c_int32), ("nonlinear", C.c_int32),
                ("coeff", (C.c_int16 * ALF_CHROMA_COEFF) * ALF_MAX_ALTS),
                ("clip_idx", (C.c_int16 * ALF_CHROMA_COEFF) * ALF_MAX_ALTS)]


class AlfParams(C.Structure):
    _fields_ = [("enabled", C.c_int32 * 3), ("num_luma_aps", C.c_int32),
                ("luma_aps", C.POINTER(AlfLumaAps)), ("chroma_aps", C.POINTER(AlfChromaAps)),
                ("ctu_enable", C.POINTER(C.c_uint8) * 3), ("ctu_filter_idx", C.POINTER(C.c_int16)),
                ("ctu_alt", C.POINTER(C.c_uint8) * 2), ("ccalf_enabled", C.c_int32 * 2),
                ("ccalf_coeff", ((C.c_int16 * CCALF_COEFF) * CCALF_MAX_FILTERS) * 2),
                ("ccalf_idc", C.POINTER(C.c_uint8) * 2), ("num_ctus", C.c_int32), ("ctu_clip", C.POINTER(C.c_uint8)),
                ("vb", C.POINTER(VirtualBoundaries))]


ALF_CLIP_TOP, ALF_CLIP_BOTTOM, ALF_CLIP_LEFT, ALF_CLIP_RIGHT, ALF_PAD_TL, ALF_PAD_BR = 1, 2, 4, 8, 16, 32


PlanePtrs = C.POINTER(C.c_int16) * 3
Strides = C.c_ssize_t * 3


HASH_MD5, HASH_CRC, HASH_CHECKSUM, HASH_SLOT_BYTES = 1, 2, 3, 48
BAND_HANDLE_BYTES = 160


class BandHandle(C.Structure):
    """vtmgpu_band_handle: opaque bytes (CUDA IPC handles of a rank's plane memory and flag block + its geometry)."""
    _fields_ = [("bytes", C.c_ubyte * BAND_HANDLE_BYTES)]


class HostPicture(C.Structure):
    """vtmgpu_host_picture: one picture of a host batch (vtmgpu_batch_filter)."""
    _fields_ = [("inp", PlanePtrs), ("in_stride", Strides), ("out", PlanePtrs), ("out_stride", Strides),
                ("deblock", C.POINTER(DeblockSparse)), ("sao", C.POINTER(SaoParams)), ("alf", C.POINTER(AlfParams))]

# every entry point include/vtmgpu.h declares (tests check that the built library exports all of them)
ENTRY_POINTS = [
    "vtmgpu_abi_version", "vtmgpu_abi_sizeof", "vtmgpu_upload_async", "vtmgpu_download_async", "vtmgpu_last_error", "vtmgpu_create", "vtmgpu_destroy", "vtmgpu_upload", "vtmgpu_download",
    "vtmgpu_set_deblock", "vtmgpu_set_deblock_async", "vtmgpu_set_deblock_sparse", "vtmgpu_set_deblock_units", "vtmgpu_upload_motion", "vtmgpu_get_deblock_records", "vtmgpu_set_lmcs", "vtmgpu_download_extended", "vtmgpu_host_register", "vtmgpu_host_unregister", "vtmgpu_set_sao", "vtmgpu_set_alf", "vtmgpu_set_alf_slices", "vtmgpu_sao_reconstruct", "vtmgpu_deblock", "vtmgpu_sao",
    "vtmgpu_alf", "vtmgpu_sao_alf", "vtmgpu_deblock_sao", "vtmgpu_filter", "vtmgpu_filter_async", "vtmgpu_sync", "vtmgpu_timer_start",
    "vtmgpu_timer_stop", "vtmgpu_rewind", "vtmgpu_launch_count", "vtmgpu_set_profiling", "vtmgpu_stage_ms",
    "vtmgpu_set_rows", "vtmgpu_set_stream", "vtmgpu_upload_rows", "vtmgpu_download_rows", "vtmgpu_export_rows", "vtmgpu_import_rows", "vtmgpu_export_halo", "vtmgpu_import_halo",
    "vtmgpu_hash",
    "vtmgpu_band_export", "vtmgpu_band_connect", "vtmgpu_band_filter_async", "vtmgpu_band_disconnect",
    "vtmgpu_batch_create", "vtmgpu_batch_destroy", "vtmgpu_batch_filter", "vtmgpu_batch_last_error", "vtmgpu_batch_launch_count",
]


def chroma_shifts(chroma_format):
    """(sx, sy) of the chroma planes for ChromaFormat 0..3 (400, 420, 422, 444)."""
    return (1 if chroma_format in (1, 2) else 0, 1 if chroma_format == 1 else 0)
