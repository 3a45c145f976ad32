"""Decoded-picture hash on the device (vtmgpu_hash): MD5 / CRC / checksum of the decoded picture hash SEI.

CPU part: the numpy restatement below (the checker of the GPU tests) against the reference's OWN calcMD5 / calcCRC /
calcChecksum (oracle/_ref/hash_ref, built from CommonLib/PicYuvMD5.cpp by oracle/Makefile) and against hashlib.
GPU part: the device digests against that restatement, hashlib, the reference binary and the MD5 the reference encoder
wrote into the SEI of every golden fixture."""
import hashlib
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden
from vvc_b200 import abi, gpu, synth

HASH_REF = os.path.join(ROOT, "oracle", "_ref", "hash_ref")


def md5_planes(planes, bds):
    """calcMD5, PicYuvMD5.cpp:188-212: per component, raster order, one byte per sample up to 8 bits, else two (little endian)."""
    return b"".join(hashlib.md5(p.astype("<u2").tobytes() if bd > 8 else p.astype(np.uint8).tobytes()).digest() for p, bd in zip(planes, bds))


def crc_planes(planes, bds):
    """calcCRC / compCRC, PicYuvMD5.cpp:88-142: CRC-16 0x1021, initial value 0xffff, message bits shifted in at the bottom, per
    sample the low byte then (bit depth > 8) the high byte, MSB first, 16 flush bits."""
    out = b""
    for p, bd in zip(planes, bds):
        v = p.astype(np.uint16).reshape(-1)
        data = np.stack([v & 0xff, v >> 8], axis=1).reshape(-1) if bd > 8 else (v & 0xff)
        crc = 0xffff
        for byte in data.astype(np.uint8).tolist():
            for b in range(7, -1, -1):
                msb = (crc >> 15) & 1
                crc = (((crc << 1) + ((byte >> b) & 1)) & 0xffff) ^ (msb * 0x1021)
        for _ in range(16):
            msb = (crc >> 15) & 1
            crc = ((crc << 1) & 0xffff) ^ (msb * 0x1021)
        out += bytes([crc >> 8, crc & 0xff])
    return out


def checksum_planes(planes, bds):
    """calcChecksum / compChecksum, PicYuvMD5.cpp:144-186."""
    out = b""
    for p, bd in zip(planes, bds):
        h, w = p.shape
        x, y = np.meshgrid(np.arange(w, dtype=np.int64), np.arange(h, dtype=np.int64))
        mask = ((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8)) & 0xff
        v = p.astype(np.int64) & 0xffff
        s = int(((v & 0xff) ^ mask).sum())
        if bd > 8:
            s += int(((v >> 8) ^ mask).sum())
        out += struct.pack(">I", s & 0xffffffff)
    return out


def reference_hashes(planes, seq, tmp_path):
    """The reference's own functions on these planes (oracle/_ref/hash_ref)."""
    f = tmp_path / "planes.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<5i", seq["width"], seq["height"], seq["chroma_format"], seq["bit_depth_luma"], seq["bit_depth_chroma"]))
        for p in planes:
            fh.write(np.ascontiguousarray(p).astype("<i2").tobytes())
    r = subprocess.run([HASH_REF, str(f)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return {ln.split()[0]: bytes.fromhex(ln.split()[1]) for ln in r.stdout.strip().splitlines()}


def _bds(seq, ncomp):
    return [seq["bit_depth_luma"]] + [seq["bit_depth_chroma"]] * (ncomp - 1)


CASES = [(64, 32, 1, 10), (72, 40, 1, 10), (200, 8, 3, 12), (136, 72, 2, 10), (320, 192, 0, 8), (384, 256, 1, 8)]


@pytest.mark.parametrize("w,h,cf,bd", CASES)
def test_restatement_equals_reference(w, h, cf, bd, tmp_path):
    if not os.path.exists(HASH_REF):
        pytest.skip("oracle/_ref/hash_ref not built (needs the reference sources at build time)")
    cap = synth.make_picture(w, h, chroma_format=cf, bit_depth=bd, seed=w + h)
    ref = reference_hashes(cap.pre, cap.seq, tmp_path)
    bds = _bds(cap.seq, cap.ncomp)
    assert md5_planes(cap.pre, bds) == ref["md5"]
    assert crc_planes(cap.pre, bds) == ref["crc"]
    assert checksum_planes(cap.pre, bds) == ref["checksum"]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,cf,bd", CASES + [(1920, 1080, 1, 10)])
def test_device_hash_equals_reference(w, h, cf, bd, tmp_path):
    cap = synth.make_picture(w, h, chroma_format=cf, bit_depth=bd, seed=w + h)
    ctx = gpu.Context(cap.seq, capacity=2)
    ctx.upload(0, cap.pre)
    ctx.upload(1, [np.ascontiguousarray(p[::-1]) for p in cap.pre])        # a second slot with other content: one call, two digests
    bds = _bds(cap.seq, cap.ncomp)
    flipped = [np.ascontiguousarray(p[::-1]) for p in cap.pre]
    small = w * h <= 1 << 17
    for kind, fn in ((abi.HASH_MD5, md5_planes), (abi.HASH_CRC, crc_planes), (abi.HASH_CHECKSUM, checksum_planes)):
        if kind == abi.HASH_CRC and not small:
            continue                                                       # the bit-serial Python CRC is too slow for 1080p: the reference binary covers it below
        got = ctx.hash(0, 2, kind)
        assert got[0] == fn(cap.pre, bds) and got[1] == fn(flipped, bds), "kind %d" % kind
    if os.path.exists(HASH_REF):
        ref = reference_hashes(cap.pre, cap.seq, tmp_path)
        assert ctx.hash(0, 1, abi.HASH_MD5)[0] == ref["md5"]
        assert ctx.hash(0, 1, abi.HASH_CRC)[0] == ref["crc"]
        assert ctx.hash(0, 1, abi.HASH_CHECKSUM)[0] == ref["checksum"]
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names())
def test_device_md5_of_filtered_fixture_equals_sei(name, manifest):
    """Every golden fixture: filter on the device, hash on the device -- the MD5 equals the SEI the reference encoder wrote, and
    CRC / checksum equal the restatement on the downloaded planes; nothing but 48 bytes per hash leaves HBM."""
    cap = load_golden(name)
    ctx = gpu.Context(cap.seq)
    ctx.set_capture(0, cap)
    ctx.filter(0, 1)
    md5 = ctx.hash(0, 1, abi.HASH_MD5)[0]
    assert [md5[16 * k:16 * k + 16].hex() for k in range(cap.ncomp)] == manifest[name]["sei_md5"]
    out = ctx.download(0)
    bds = _bds(cap.seq, cap.ncomp)
    assert ctx.hash(0, 1, abi.HASH_CRC)[0] == crc_planes(out, bds)
    assert ctx.hash(0, 1, abi.HASH_CHECKSUM)[0] == checksum_planes(out, bds)
    ctx.close()


@pytest.mark.gpu
def test_hash_bad_arguments():
    cap = synth.make_picture(64, 32, seed=1)
    ctx = gpu.Context(cap.seq)
    with pytest.raises(gpu.VtmGpuError):
        ctx.hash(0, 2, abi.HASH_MD5)
    with pytest.raises(gpu.VtmGpuError):
        ctx.hash(0, 1, 9)
    ctx.close()
