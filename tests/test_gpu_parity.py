"""GPU (-m gpu): the CUDA path, called through the C ABI, against (a) the reference's stage outputs in the golden
fixtures, (b) the SEI MD5 of the reference encoder, (c) the oracle on seeded synthetic pictures of other sizes,
chroma formats and bit depths.  Bar: bit-exact (integer sample work)."""
import numpy as np
import pytest

import pyoracle
from conftest import golden_names, load_golden, plane_md5
from vvc_b200 import gpu, synth

pytestmark = pytest.mark.gpu
NAMES = golden_names()


def _eq(a, b, what):
    for c in range(len(a)):
        if not np.array_equal(a[c], b[c]):
            bad = np.argwhere(a[c] != b[c])
            raise AssertionError("%s: component %d differs at %d samples, first (y,x)=%s got %d want %d" %
                                 (what, c, len(bad), tuple(bad[0]), a[c][tuple(bad[0])], b[c][tuple(bad[0])]))


@pytest.mark.parametrize("name", NAMES)
def test_golden_staged(name):
    """loopFilterPic / SAOProcess / ALFProcess one by one, compared with the reference after every stage."""
    cap = load_golden(name)
    out = gpu.execute_loop_filters(cap, fused=False)
    _eq(out["dbf"], cap.stage["dbf"], "deblocking")
    if cap.stage["sao"] is not None:
        _eq(out["sao"], cap.stage["sao"], "SAO")
    if cap.stage["alf"] is not None:
        _eq(out["alf"], cap.stage["alf"], "ALF")


@pytest.mark.parametrize("name", NAMES)
def test_golden_fused_md5(name, manifest):
    """whole chain with SAO fused into the ALF pass == decoded picture hash of the reference encoder"""
    cap = load_golden(name)
    out = gpu.execute_loop_filters(cap, fused=True)["final"]
    assert [plane_md5(p) for p in out] == manifest[name]["sei_md5"]


SYNTH = [
    # width, height, chroma_format, bit depth, ctu, density
    (128, 128, 1, 10, 128, 1.0),      # a single CTU
    (136, 72, 1, 10, 128, 0.7),       # ragged: smaller than one tile in both directions
    (416, 240, 1, 10, 128, 0.6),
    (456, 264, 3, 10, 128, 0.8),      # 4:4:4, partial CTUs / partial tiles
    (448, 256, 2, 10, 128, 0.8),      # 4:2:2
    (320, 192, 0, 10, 128, 0.8),      # 4:0:0
    (384, 256, 1, 8, 128, 0.8),       # 8-bit
    (384, 256, 1, 12, 128, 0.8),      # 12-bit
    (384, 320, 1, 10, 64, 0.8),       # CTU 64
    (416, 240, 1, 10, 32, 0.8),       # CTU 32: 2 x 2 CTUs per ALF tile, partial tiles
    (328, 200, 3, 10, 32, 1.0),       # CTU 32, 4:4:4, ragged
    (1920, 1080, 1, 10, 128, 0.5),
    # degenerate geometries: pictures smaller than one tile / one TMA box, single rows of units
    (8, 8, 1, 10, 128, 1.0),
    (16, 8, 1, 10, 128, 1.0),
    (72, 40, 1, 10, 128, 1.0),
    (24, 136, 3, 10, 128, 1.0),
    (64, 64, 2, 10, 64, 1.0),
    (200, 8, 1, 10, 128, 1.0),
]


@pytest.mark.parametrize("w,h,cf,bd,ctu,density", SYNTH)
def test_synthetic_vs_oracle(w, h, cf, bd, ctu, density):
    cap = synth.make_picture(w, h, chroma_format=cf, bit_depth=bd, ctu_size=ctu, seed=w + h + cf, density=density)
    want = pyoracle.filter_capture(cap)
    got = gpu.execute_loop_filters(cap, fused=False)
    for st in ("dbf", "sao", "alf"):
        _eq(got[st], want[st], st)
    fused = gpu.execute_loop_filters(cap, fused=True)["final"]
    _eq(fused, want["final"], "fused chain")


@pytest.mark.parametrize("w,h,cf,ctu", [(512, 384, 1, 128), (456, 264, 3, 128), (448, 256, 2, 64), (320, 320, 0, 64), (1920, 1080, 1, 128),
                                         (416, 240, 1, 32), (328, 200, 3, 32)])
def test_partition_boundaries_vs_oracle(w, h, cf, ctu):
    """ALF at slice / tile boundaries that must not be crossed: random per-CTU clip and corner-pad flags,
    forced-on ALF / CC-ALF; the fixtures tiles_* / slices_* pin the same path against the reference itself."""
    cap = synth.make_picture(w, h, chroma_format=cf, ctu_size=ctu, seed=3 * w + cf, density=1.0, partitions=True)
    assert cap.alf["clip"].any()
    want = pyoracle.filter_capture(cap)
    got = gpu.execute_loop_filters(cap, fused=False)
    _eq(got["alf"], want["alf"], "ALF with partition boundaries")
    _eq(gpu.execute_loop_filters(cap, fused=True)["final"], want["final"], "fused chain")


@pytest.mark.parametrize("w,h,cf,bd", [(512, 384, 1, 10), (456, 264, 3, 8), (448, 256, 2, 12), (1920, 1080, 1, 10)])
def test_ladf_vs_oracle(w, h, cf, bd):
    """LADF (deriveLADFShift, LoopFilter.cpp:815-841): records carry QPs, tc / beta are derived on the device from the
    reconstructed samples next to the edge (after the vertical pass for horizontal edges); dense and list form."""
    cap = synth.make_picture(w, h, chroma_format=cf, bit_depth=bd, seed=w + bd, density=0.9, ladf=True)
    want = pyoracle.filter_capture(cap)
    got = gpu.execute_loop_filters(cap, fused=False)
    _eq(got["dbf"], want["dbf"], "deblocking with LADF")
    ctx = gpu.Context(cap.seq)
    ctx.set_capture(0, cap)
    ctx.set_deblock_sparse(0, gpu.sparse_records(cap.dbf_luma, cap.dbf_chroma if cap.ncomp > 1 else None, ladf=cap.ladf_struct()))
    ctx.filter(0, 1)
    _eq(ctx.download(0), want["final"], "LADF, records as lists")
    ctx.close()


VB_CASES = [
    # width, height, chroma_format, ctu, vertical boundaries, horizontal boundaries, partitions
    (512, 384, 1, 128, [200, 384], [248], False),         # inside a CTU / on a CTU edge; 4 rows above the ALF boundary of the CTU row
    (512, 384, 1, 128, [64, 256], [128, 320], True),      # on tile edges inside a CTU, on CTU edges, combined with slice / tile clipping
    (456, 264, 3, 128, [8, 448], [256], False),           # 4:4:4, next to the picture border, last (partial) tile
    (448, 256, 2, 64, [40, 104, 232], [24, 120, 184], False),   # 4:2:2, CTU 64, three per direction, 2 x 2 parts in one tile
    (320, 320, 0, 64, [160], [], False),                  # 4:0:0
    (1920, 1080, 1, 128, [960], [544], False),
    (448, 256, 1, 32, [40, 104, 168], [24, 56, 200], True),     # CTU 32: CTU cuts and virtual-boundary cuts in one tile
    (416, 240, 1, 32, [32, 64, 224], [96], False),              # CTU 32: boundaries on CTU edges
]


@pytest.mark.parametrize("w,h,cf,ctu,vx,vy,parts", VB_CASES)
def test_virtual_boundaries_vs_oracle(w, h, cf, ctu, vx, vy, parts):
    """Signalled virtual boundaries (picture header): SAO leaves the samples next to them alone for the edge classes that look
    across, ALF pads every part of a CTU between boundaries separately (tiles cut by a boundary are filtered part by part)."""
    cap = synth.make_picture(w, h, chroma_format=cf, ctu_size=ctu, seed=5 * w + cf, density=1.0, partitions=parts, vb=(vx, vy))
    want = pyoracle.filter_capture(cap)
    got = gpu.execute_loop_filters(cap, fused=False)
    _eq(got["sao"], want["sao"], "SAO with virtual boundaries")
    _eq(got["alf"], want["alf"], "ALF with virtual boundaries")
    _eq(gpu.execute_loop_filters(cap, fused=True)["final"], want["final"], "fused chain")


def test_stage_switches():
    """NULL side info switches a stage off: the picture must pass through unchanged."""
    cap = synth.make_picture(256, 128, seed=5)
    ctx = gpu.Context(cap.seq)
    ctx.upload(0, cap.pre)
    ctx.set_deblock(0, None)
    ctx.set_sao(0, None)
    ctx.set_alf(0, None)
    ctx.filter(0, 1)
    _eq(ctx.download(0), cap.pre, "all stages off")
    ctx.close()


def test_batch_of_slots_and_rewind():
    """picture-parallel: several independent pictures in one launch sequence; rewind repeats the result"""
    caps = [synth.make_picture(384, 256, seed=s, density=0.7) for s in range(5)]
    want = [pyoracle.filter_capture(c)["final"] for c in caps]
    ctx = gpu.Context(caps[0].seq, capacity=len(caps))
    for s, c in enumerate(caps):
        ctx.set_capture(s, c)
    for rep in range(2):
        ctx.filter(0, len(caps))
        for s in range(len(caps)):
            _eq(ctx.download(s), want[s], "slot %d rep %d" % (s, rep))
        ctx.rewind(0, len(caps))
    assert ctx.launch_count() == 4 + len(caps)       # two chain kernels per repetition for the whole batch + one queue build per picture when its records were set
    ctx.close()


def test_full_size_4k_properties():
    """3840x2160 (BASELINE config 3 size): GPU == oracle on one forced-on picture, and the chain is deterministic."""
    cap = synth.make_picture(3840, 2160, seed=2160, density=1.0)
    ctx = gpu.Context(cap.seq)
    ctx.set_capture(0, cap)
    ctx.filter(0, 1)
    a = ctx.download(0)
    ctx.rewind(0, 1)
    ctx.filter(0, 1)
    b = ctx.download(0)
    ctx.close()
    _eq(a, b, "determinism")
    _eq(a, pyoracle.filter_capture(cap)["final"], "4K forced-on vs oracle")


def test_async_side_info_and_caller_stream():
    """vtmgpu_set_deblock_async (no staging copy) and vtmgpu_set_stream (all work enqueued on the caller's stream) give
    the same picture as the synchronous path."""
    import torch
    cap = synth.make_picture(640, 384, seed=21, density=0.8)
    want = pyoracle.filter_capture(cap)["final"]
    ctx = gpu.Context(cap.seq)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream, True)
    ctx.upload(0, cap.pre, sync=False)
    dp = cap.deblock_params()                       # stays alive until the sync below
    ctx.set_deblock(0, dp, sync=False)
    ctus = cap.sao_ctus()
    gpu.sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, cap.sao_scale[0], cap.sao_scale[1])
    ctx.set_sao(0, ctus)
    ctx.set_alf(0, cap.alf_params())
    ctx.filter(0, 1)                                # only enqueues in this mode
    stream.synchronize()
    _eq(ctx.download(0), want, "caller-stream chain")
    ctx.set_stream(None, False)
    ctx.close()


@pytest.mark.parametrize("w,h,cf,density", [(640, 384, 1, 0.3), (456, 264, 3, 1.0), (320, 192, 0, 0.1), (1920, 1080, 1, 0.05)])
def test_sparse_records_equal_dense(w, h, cf, density):
    """vtmgpu_set_deblock_sparse (lists of the active units, scattered on the device) == the dense record arrays; a slot that
    held a denser picture before must not keep stale records; records listed on the picture border are ignored."""
    dense_before = synth.make_picture(w, h, chroma_format=cf, seed=77, density=1.0)
    cap = synth.make_picture(w, h, chroma_format=cf, seed=w + cf, density=density)
    want = pyoracle.filter_capture(cap)["final"]
    ctx = gpu.Context(cap.seq)
    ctx.set_capture(0, dense_before)
    ctx.filter(0, 1)
    ctx.rewind(0, 1)
    ctx.set_capture(0, cap)
    sp = gpu.sparse_records(cap.dbf_luma, cap.dbf_chroma if cap.ncomp > 1 else None, pin=(cf == 1))
    assert sp.luma_count[0] <= cap.dbf_luma[0].size
    ctx.set_deblock_sparse(0, sp)
    ctx.filter(0, 1)
    _eq(ctx.download(0), want, "sparse records")
    ctx.close()


@pytest.mark.parametrize("lanes", [1, 2, 3, 8])
def test_host_batch_equals_oracle(lanes):
    """vtmgpu_batch_filter: a run of host pictures through the whole boundary in one C call (round robin over the lanes, more
    pictures than lanes, pictures with stages switched off, in-place output) -- every output equals the oracle's.  One and two
    lanes take the direct download, three and more the download that waits for the next picture's upload; eight lanes > pictures."""
    caps = [synth.make_picture(448, 256, chroma_format=1, seed=40 + i, density=0.3 + 0.1 * i) for i in range(7)]
    want = [pyoracle.filter_capture(c)["final"] for c in caps]
    pre0 = [[p.copy() for p in c.pre] for c in caps]              # picture 2 is filtered in place below
    batch = gpu.Batch(caps[0].seq, lanes=lanes)
    pics, outs = [], []
    for i, c in enumerate(caps):
        ctus = c.sao_ctus()
        gpu.sao_reconstruct(ctus, c.width_in_ctus, c.ncomp, c.sao_scale[0], c.sao_scale[1])
        inp = [np.ascontiguousarray(p) for p in c.pre]
        out = inp if i == 2 else [np.zeros_like(p) for p in inp]                 # picture 2 is filtered in place
        outs.append(out)
        pics.append(gpu.host_picture(inp, out, gpu.sparse_records(c.dbf_luma, c.dbf_chroma), ctus, c.alf_params()))
    batch.filter(pics)
    for i in range(len(caps)):
        _eq(outs[i], want[i], "host batch picture %d" % i)
    # all stages off: the pictures come back unchanged
    c = caps[0]
    inp = [np.ascontiguousarray(p) for p in c.pre]
    out = [np.zeros_like(p) for p in inp]
    batch.filter([gpu.host_picture(inp, out)])
    _eq(out, c.pre, "host batch, stages off")
    # the same batch object again (the lanes' events of the previous call are all complete), page-locked buffers this time
    import torch
    pin = [[torch.from_numpy(p).pin_memory() for p in pl] for pl in pre0]
    pout = [[torch.zeros_like(t).pin_memory() for t in pl] for pl in pin]
    pics = []
    for i, c in enumerate(caps):
        ctus = c.sao_ctus()
        gpu.sao_reconstruct(ctus, c.width_in_ctus, c.ncomp, c.sao_scale[0], c.sao_scale[1])
        pics.append(gpu.host_picture([t.numpy() for t in pin[i]], [t.numpy() for t in pout[i]], gpu.sparse_records(c.dbf_luma, c.dbf_chroma, pin=True), ctus, c.alf_params()))
    for _ in range(2):
        batch.filter(pics)
    for i in range(len(caps)):
        _eq([t.numpy() for t in pout[i]], want[i], "host batch (page-locked) picture %d" % i)
    assert batch.launch_count() > 0
    batch.close()


def test_bad_arguments_fail_loudly():
    cap = synth.make_picture(256, 128, seed=1)
    ctx = gpu.Context(cap.seq)
    with pytest.raises(gpu.VtmGpuError):
        ctx.filter(0, 2)                       # capacity is 1
    p = cap.alf_params()
    p.num_ctus += 1
    with pytest.raises(gpu.VtmGpuError):
        ctx.set_alf(0, p)
    sp = gpu.sparse_records(cap.dbf_luma, cap.dbf_chroma)
    sp.luma_count[0] = cap.dbf_luma[0].size + 1
    with pytest.raises(gpu.VtmGpuError):
        ctx.set_deblock_sparse(0, sp)
    # band-mode row / halo copies: slot index and buffers are validated before anything is indexed or enqueued
    import torch
    buf = torch.zeros(4 * 3 * 256, dtype=torch.int16, device="cuda")
    for bad_slot in (-1, 1, 1 << 20):
        with pytest.raises(gpu.VtmGpuError):
            ctx.export_halo(bad_slot, [0, 0, 0], 4, buf.data_ptr())
        with pytest.raises(gpu.VtmGpuError):
            ctx.import_halo(bad_slot, [0, 0, 0], 4, buf.data_ptr())
        with pytest.raises(gpu.VtmGpuError):
            ctx.export_rows(bad_slot, 0, 0, 4, buf.data_ptr())
        with pytest.raises(gpu.VtmGpuError):
            ctx.import_rows(bad_slot, 0, 0, 4, buf.data_ptr())
    with pytest.raises(gpu.VtmGpuError):
        ctx.export_halo(0, [0, 0, 0], 4, None)
    with pytest.raises(gpu.VtmGpuError):
        ctx.import_halo(0, [0, 0, 0], 4, None)
    ctx.export_halo(0, [0, 0, 0], 4, buf.data_ptr())          # and the good call still works
    ctx.sync()
    ctx.close()


STREAMS = [("ra_416x240.bin", 8), ("ld444_1080p.bin", 16), ("ra_1080p.bin", 32), ("ai_4320p.bin", 1),
           ("ra_2160p_8.bin", 8), ("ra_2160p_b.bin", 32),     # BASELINE config 3: 3840x2160 RA with CC-ALF (the benchmark content: frames 0..7 and 32..63)
           # tiles / raster-scan slices with in-loop filtering across their boundaries disabled
           ("tiles_832x480.bin", 5), ("slices_832x480.bin", 5), ("slices45_832x480.bin", 3),
           ("ladf_832x480.bin", 5),       # LADF: deblocking thresholds derived on the device
           ("vb_832x480.bin", 5),         # signalled virtual boundaries inside CTUs and on CTU edges
           ("ld422_416x240.bin", 4),      # 4:2:2, full CTC tool set
           ("ctu64_416x240.bin", 4), ("ctu32_416x240.bin", 4), ("bd12_416x240.bin", 3), ("dbfoffs_416x240.bin", 4),   # CTU 64, 12-bit, beta / tc offsets
           ("scc444_416x240.bin", 3), ("ldp_416x240.bin", 4),
           ("lmcs_416x240.bin", 6),       # LMCS with the slice reshaper on: inverse luma mapping folded into the deblocking kernel's tile load
           ("ra_full_832x480.bin", 6)]    # full CTC tool set at a size with 128-wide CUs   # palette / IBC / BDPCM on screen content (4:4:4); P slices


def _fuzz_streams():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streams", "fuzz_manifest.json")
    if not os.path.exists(path):
        return []
    with open(path) as f:
        return [(name, m["pictures"]) for name, m in sorted(json.load(f).items())]


# random encoder configurations (tools/fuzz_parity_streams.py; options in tests/golden/streams/fuzz_manifest.json): combinations of the
# path features above that no hand-picked stream has -- e.g. 4:4:4 + raster slices + virtual boundaries, 4:2:2 + CTU 32 + LADF
STREAMS += _fuzz_streams()


@pytest.mark.parametrize("stream,pictures", STREAMS)
def test_decoder_drop_in_md5(stream, pictures):
    """The reference decoder with OUR filter entry points linked in (vvc_b200/_bin/DecoderApp_gpu) decodes reference-encoded
    streams -- full CTC tool set RA 416x240, 4:4:4 low delay 1080p with chroma ALF + CC-ALF (BASELINE config 5), RA 1080p
    (config 2), one 8K intra picture (config 4) -- and every picture matches the MD5 SEI the reference encoder wrote."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dec = os.path.join(root, "vvc_b200", "_bin", "DecoderApp_gpu")
    if not os.path.exists(dec):
        pytest.skip("DecoderApp_gpu not built (needs the reference sources at build time)")
    r = subprocess.run([dec, "-b", os.path.join(root, "tests", "golden", "streams", stream), "-d", "0"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, VTMGPU_SHIM_BACKEND="gpu"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("(OK)") == pictures and "ERROR" not in r.stdout, r.stdout[-2000:]


YUV_STREAMS = ["ra_416x240.bin", "ra_1080p.bin", "ld444_1080p.bin", "ra_2160p_8.bin", "lmcs_416x240.bin"]      # BASELINE configs 1, 2, 5, 3


@pytest.mark.parametrize("stream", YUV_STREAMS)
def test_decoder_output_yuv_byte_identical(stream, tmp_path):
    """north_star: "byte-identical output YUV".  The reconstructed YUV written by the decoder with OUR filters (DecoderApp_gpu -o)
    equals, byte for byte, the file the UNMODIFIED reference decoder (oracle/_ref/DecoderApp -o, its own CPU filters) writes for the
    same stream -- this also covers what the MD5 SEI does not (output order, bit depth conversion, cropping of the written file)."""
    import filecmp
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dec, ref = os.path.join(root, "vvc_b200", "_bin", "DecoderApp_gpu"), os.path.join(root, "oracle", "_ref", "DecoderApp")
    if not (os.path.exists(dec) and os.path.exists(ref)):
        pytest.skip("DecoderApp_gpu / the reference DecoderApp are not built (they need the reference sources at build time)")
    bits = os.path.join(root, "tests", "golden", "streams", stream)
    ours, theirs = str(tmp_path / "gpu.yuv"), str(tmp_path / "ref.yuv")
    r = subprocess.run([dec, "-b", bits, "-o", ours, "-d", "0"], capture_output=True, text=True, timeout=900, env=dict(os.environ, VTMGPU_SHIM_BACKEND="gpu"))
    assert r.returncode == 0 and "ERROR" not in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([ref, "-b", bits, "-o", theirs, "-d", "0"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ERROR" not in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.getsize(ours) > 0 and os.path.getsize(ours) == os.path.getsize(theirs)
    assert filecmp.cmp(ours, theirs, shallow=False), "%s: the GPU decoder's YUV differs from the reference decoder's" % stream


@pytest.mark.parametrize("stream,pictures,lmcs", [("lmcs_416x240.bin", 6, True), ("ra_416x240.bin", 8, False), ("ra_1080p.bin", 32, False)])
def test_decoder_host_passes_on_device(stream, pictures, lmcs):
    """SURVEY 8f n2 at decoder level.  With the slice reshaper on (lmcs_*: PQ signal type; the encoder's SDR analysis leaves it off on the
    other synthetic clips) executeLoopFilters maps the luma reconstruction through the inverse table on the host (DecLib.cpp:570-577),
    and the border of every reference picture is extended on the host (Picture.cpp:737).  In DecoderApp_gpu both happen on the device -- the host mapping is bypassed (k_dbf_sao maps while it loads its
    tiles) and the margins arrive with the download; the inter-predicted pictures that follow read those margins, so `MD5 (OK)` on
    every picture covers both.  The shim's counters prove the device paths were the ones that ran; with both switched back to the
    host the stream decodes to the same MD5s."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dec = os.path.join(root, "vvc_b200", "_bin", "DecoderApp_gpu")
    if not os.path.exists(dec):
        pytest.skip("DecoderApp_gpu not built (needs the reference sources at build time)")
    bits = os.path.join(root, "tests", "golden", "streams", stream)

    def run(**env):
        r = subprocess.run([dec, "-b", bits, "-d", "0"], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, VTMGPU_SHIM_BACKEND="gpu", VTMGPU_SHIM_TIMING="1", **env))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert r.stdout.count("(OK)") == pictures and "ERROR" not in r.stdout, r.stdout[-2000:]
        m = re.search(r"lmcs_on_device=(\d+) border_on_device=(\d+)", r.stdout)
        assert m, r.stdout[-500:]
        return int(m.group(1)), int(m.group(2)), re.findall(r"\[MD5:[0-9a-f,]+", r.stdout)

    lm, bo, md5 = run()
    assert lm == (pictures if lmcs else 0), "%s: %d pictures took the device LMCS path" % (stream, lm)
    assert bo == pictures
    lm0, bo0, md5_host = run(VTMGPU_SHIM_HOST_LMCS="1", VTMGPU_SHIM_EXTEND="0", VTMGPU_SHIM_DEVICE_DERIVE="0", VTMGPU_SHIM_PIN="0")
    assert lm0 == 0 and bo0 == 0 and md5_host == md5


@pytest.mark.parametrize("stream,pictures", STREAMS)
def test_decoder_device_derivation(stream, pictures):
    """with VTMGPU_SHIM_DEVICE_DERIVE=1 the shim sends the flattened block structure (CUs, TUs, unit maps, slices, the motion
    field) and k_dbf_derive produces the deblocking records on the device.  VTMGPU_SHIM_CHECK_UNITS=1 makes the shim run the CU walk as
    well and compare, for every picture, (a) the derivation source run on the host and (b) the records the KERNEL wrote
    (vtmgpu_get_deblock_records) with the walk's arrays -- byte-identical -- and every picture must still decode to `MD5 (OK)`."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dec = os.path.join(root, "vvc_b200", "_bin", "DecoderApp_gpu")
    if not os.path.exists(dec):
        pytest.skip("DecoderApp_gpu not built (needs the reference sources at build time)")
    r = subprocess.run([dec, "-b", os.path.join(root, "tests", "golden", "streams", stream), "-d", "0"], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, VTMGPU_SHIM_BACKEND="gpu", VTMGPU_SHIM_DEVICE_DERIVE="1", VTMGPU_SHIM_CHECK_UNITS="1", VTMGPU_SHIM_TIMING="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("(OK)") == pictures and "ERROR" not in r.stdout, r.stdout[-2000:]
    m = re.search(r"device_derived=(\d+) walk_derived_instead=(\d+) units_checked=(\d+) kernel_records_checked=(\d+)", r.stdout)
    assert m, r.stdout[-500:]
    assert [int(v) for v in m.groups()] == [pictures, 0, pictures, pictures]
