"""CPU: the C-ABI library loads, exports every symbol include/vtmgpu.h declares, the ctypes mirrors have the
compiled sizes, and the host-only helper (SAO parameter reconstruction) agrees with the oracle.  No kernel runs."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pyoracle
from conftest import ROOT
from vvc_b200 import abi, gpu, synth


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(gpu.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return gpu.load_library()


def test_header_and_binding_agree():
    hdr = open(os.path.join(ROOT, "include", "vtmgpu.h")).read()
    declared = set(re.findall(r"\b(vtmgpu_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(abi.ENTRY_POINTS)


def test_exports_every_declared_symbol(lib):
    for name in abi.ENTRY_POINTS:
        assert hasattr(lib, name), name
    assert lib.vtmgpu_abi_version() == abi.ABI_VERSION


def test_struct_sizes(lib):
    mirrors = [abi.SeqParams, abi.DeblockParams, abi.SaoOffset, abi.SaoCtu, abi.SaoParams, abi.AlfLumaAps, abi.AlfChromaAps, abi.AlfParams, abi.DeblockSparse, abi.Ladf, abi.VirtualBoundaries, abi.HostPicture]
    for i, m in enumerate(mirrors):
        assert lib.vtmgpu_abi_sizeof(i) == C.sizeof(m), m.__name__


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(gpu.VtmGpuError):
        gpu.load_library(str(tmp_path / "libvtmgpu.so"))


def test_no_cpu_fallback_without_device(lib):
    """On a box without a CUDA device vtmgpu_create must fail (on the GPU box it succeeds: then just destroy)."""
    sp = abi.SeqParams(64, 64, 1, 10, 10, 128, 1, 0)
    h = C.c_void_p()
    rc = lib.vtmgpu_create(C.byref(sp), C.byref(h))
    if rc == 0:
        lib.vtmgpu_destroy(h)
    else:
        assert b"no CUDA device" in lib.vtmgpu_last_error(None) or b"CUDA" in lib.vtmgpu_last_error(None)


@pytest.mark.parametrize("bad", [dict(width=100), dict(height=0), dict(chroma_format=4), dict(bit_depth_luma=7), dict(ctu_size=16), dict(ctu_size=256), dict(capacity=0)])
def test_create_rejects_bad_geometry(lib, bad):
    kw = dict(width=64, height=64, chroma_format=1, bit_depth_luma=10, bit_depth_chroma=10, ctu_size=128, capacity=1, device=0)
    kw.update(bad)
    sp = abi.SeqParams(*[kw[k] for k, _ in abi.SeqParams._fields_])
    h = C.c_void_p()
    assert lib.vtmgpu_create(C.byref(sp), C.byref(h)) != 0
    assert lib.vtmgpu_last_error(None)


@pytest.mark.parametrize("seed", range(6))
def test_sao_reconstruct_matches_oracle(lib, seed):
    cap = synth.make_picture(640, 384, chroma_format=1, seed=seed, density=0.9)
    a, b = cap.sao_ctus(), cap.sao_ctus()
    ma = gpu.sao_reconstruct(a, cap.width_in_ctus, cap.ncomp, 0, seed % 2)
    mb = pyoracle.sao_reconstruct(b, cap.width_in_ctus, cap.ncomp, 0, seed % 2)
    assert ma == mb and bytes(a) == bytes(b)
    assert all(a[i].comp[c].mode != abi.SAO_MODE_MERGE for i in range(len(a)) for c in range(3))


def test_sao_reconstruct_rejects_missing_merge_target(lib):
    ctus = (abi.SaoCtu * 2)()
    ctus[0].comp[0].mode = abi.SAO_MODE_MERGE
    ctus[0].comp[0].type = abi.SAO_MERGE_LEFT      # CTU 0 has no left neighbour
    with pytest.raises(gpu.VtmGpuError):
        gpu.sao_reconstruct(ctus, 2, 3)


def test_sparse_records_layout():
    """gpu.sparse_records: lists of the non-zero records, in ONE buffer in array order, each list on the next 16-byte boundary
    (the layout for which vtmgpu_set_deblock_sparse needs a single upload)."""
    import numpy as np
    from vvc_b200 import synth
    cap = synth.make_picture(256, 128, seed=3, density=0.4)
    sp = gpu.sparse_records(cap.dbf_luma, cap.dbf_chroma)
    base = C.cast(sp.luma[0], C.c_void_p).value
    off = 0
    for d in range(2):
        n = int(np.count_nonzero(cap.dbf_luma[d]))
        assert sp.luma_count[d] == n
        assert C.cast(sp.luma[d], C.c_void_p).value == base + off
        ent = np.ctypeslib.as_array(C.cast(sp.luma[d], C.POINTER(C.c_uint32)), shape=(n, 2))
        assert np.array_equal(cap.dbf_luma[d][ent[:, 0]], ent[:, 1]) and len(set(ent[:, 0].tolist())) == n
        off = (off + n * C.sizeof(abi.DbfLumaEntry) + 15) & ~15
    for d in range(2):
        n = int(np.count_nonzero(cap.dbf_chroma[d]))
        assert sp.chroma_count[d] == n
        assert C.cast(sp.chroma[d], C.c_void_p).value == base + off
        off = (off + n * C.sizeof(abi.DbfChromaEntry) + 15) & ~15
    assert base % 16 == 0 and sp.nbytes == off


def test_header_is_plain_c(tmp_path):
    """include/vtmgpu.h is the boundary for cgo / JNI / ctypes style bindings: it must compile as C99 without C++ or CUDA headers."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "vtmgpu.h")
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
