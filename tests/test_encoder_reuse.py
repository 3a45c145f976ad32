"""SURVEY 8f row n4 -- encoder-side reuse.  The reference ENCODER calls the same LoopFilter::loopFilterPic on its final reconstruction
(EncoderLib/EncGOP.cpp:2794; :3436 for the parameter-selection variant).  vvc_b200/_bin/EncoderApp_gpu is the unmodified reference
encoder with OUR LoopFilter linked in (vvc_b200/shim/Makefile; its SAO / ALF parameter searches keep the reference's own classes, which
they derive from).  Everything downstream of the deblocked picture -- SAO statistics, ALF filter design, the reference pictures of the
following frames, the MD5 SEI -- depends on every deblocked sample, so the bar is: the bitstream and the reconstruction file are
BYTE-IDENTICAL to what the stock reference encoder (oracle/_ref/EncoderApp) writes for the same command line."""
import filecmp
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC_GPU = os.path.join(ROOT, "vvc_b200", "_bin", "EncoderApp_gpu")
ENC_REF = os.path.join(ROOT, "oracle", "_ref", "EncoderApp")
DEC_REF = os.path.join(ROOT, "oracle", "_ref", "DecoderApp")

# the reference's own low-delay configuration, copied next to the reference binaries by oracle/Makefile (oracle/_ref is git-ignored and
# travels to the GPU box with them)
CFG = os.path.join(ROOT, "oracle", "_ref", "cfg", "encoder_lowdelay_vtm.cfg")
# full CTC tool set, shallower partitioning search for speed
TOOLS = "--MaxMTTHierarchyDepth=1 --MaxMTTHierarchyDepthISliceL=1 --MaxMTTHierarchyDepthISliceC=1 --SearchRange=32 --MTS=0 --LFNST=0".split()


def _command(enc, tmp, tag, w, h, frames, qp, extra):
    return [enc, "-c", CFG, "-i", os.path.join(tmp, "src.yuv"), "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames),
            "--InputBitDepth=10", "--InternalBitDepth=10", "--InputChromaFormat=420", "-q", str(qp), "--SEIDecodedPictureHash=1",
            "-b", os.path.join(tmp, tag + ".bin"), "-o", os.path.join(tmp, tag + ".yuv")] + TOOLS + extra


def test_encoder_without_a_device_fails_loudly(tmp_path):
    """CPU: the GPU encoder has no CPU deblocking to fall back to -- without a CUDA device the first picture aborts the run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    if not (os.path.exists(ENC_GPU) and os.path.exists(CFG)):
        pytest.skip("EncoderApp_gpu not built (needs the reference sources at build time)")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_streams
    tmp = str(tmp_path)
    make_streams.gen_yuv(os.path.join(tmp, "src.yuv"), 64, 64, 420, 1, 3, 10)
    r = subprocess.run(_command(ENC_GPU, tmp, "gpu", 64, 64, 1, 32, []), capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and ("CUDA" in r.stdout + r.stderr or "libvtmgpu" in r.stdout + r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,frames,qp,extra", [(416, 240, 6, 32, []), (416, 240, 4, 27, ["--LoopFilterBetaOffset_div2=2", "--LoopFilterTcOffset_div2=-3", "--LADF=1"])])
def test_encoder_with_gpu_deblocking_writes_the_same_stream(tmp_path, w, h, frames, qp, extra):
    if not (os.path.exists(ENC_GPU) and os.path.exists(ENC_REF) and os.path.exists(CFG)):
        pytest.skip("EncoderApp_gpu / the reference EncoderApp are not built (they need the reference sources at build time)")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_streams
    tmp = str(tmp_path)
    make_streams.gen_yuv(os.path.join(tmp, "src.yuv"), w, h, 420, frames, 71 + qp, 14)
    gpu = subprocess.Popen(_command(ENC_GPU, tmp, "gpu", w, h, frames, qp, extra), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    ref = subprocess.run(_command(ENC_REF, tmp, "ref", w, h, frames, qp, extra), capture_output=True, text=True, timeout=900)
    out, _ = gpu.communicate(timeout=900)
    assert ref.returncode == 0, ref.stdout[-2000:]
    assert gpu.returncode == 0, out[-2000:]
    assert os.path.getsize(os.path.join(tmp, "gpu.bin")) > 0
    assert filecmp.cmp(os.path.join(tmp, "gpu.bin"), os.path.join(tmp, "ref.bin"), shallow=False), "bitstreams differ"
    assert filecmp.cmp(os.path.join(tmp, "gpu.yuv"), os.path.join(tmp, "ref.yuv"), shallow=False), "reconstructions differ"
    # and the stock decoder accepts it (MD5 SEI written from the GPU-deblocked reconstruction)
    d = subprocess.run([DEC_REF, "-b", os.path.join(tmp, "gpu.bin"), "-d", "0"], capture_output=True, text=True, timeout=300)
    assert d.returncode == 0 and "ERROR" not in d.stdout and d.stdout.count("(OK)") >= 2, d.stdout[-1000:]
