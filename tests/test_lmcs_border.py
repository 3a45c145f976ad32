"""the two whole-picture passes either side of the filter chain, on the device.
  * LMCS inverse luma mapping (AreaBuf<Pel>::rspSignal, Buffer.cpp:380-393, called at DecLib.cpp:570-577) folded into the tile load of
    the deblocking / SAO kernel (vtmgpu_set_lmcs);
  * reference-picture border extension (Picture::extendPicBorder, Picture.cpp:737-772) as part of the download (vtmgpu_download_extended).
CPU: the oracle restatements against their definitions.  GPU: through the C ABI against the oracle, bit-exact; the decoder-level check
(LMCS streams decode MD5 (OK) with the host call bypassed) is in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import pyoracle
from vvc_b200 import synth


def _lut(bd, seed):
    """a monotone piecewise-linear inverse table like Reshape::constructReshaper builds (16 pieces), any table is a legal input"""
    rng = np.random.default_rng(seed)
    n = 1 << bd
    knots = np.sort(rng.choice(np.arange(1, n - 1), size=15, replace=False))
    xs = np.concatenate([[0], knots, [n - 1]])
    ys = np.sort(rng.integers(0, n, size=17))
    return np.rint(np.interp(np.arange(n), xs, ys)).astype(np.int16)


def test_oracle_lmcs_is_a_table_lookup():
    cap = synth.make_picture(64, 64, seed=1)
    lut = _lut(10, 3)
    planes = [p.copy() for p in cap.pre]
    pyoracle.lmcs_inverse(planes, lut)
    ys, xs = np.nonzero(np.ones_like(cap.pre[0]))
    assert all(planes[0][y, x] == lut[cap.pre[0][y, x]] for y, x in zip(ys[::97], xs[::97]))
    assert np.array_equal(planes[1], cap.pre[1]) and np.array_equal(planes[2], cap.pre[2])


def test_oracle_border_extension_is_nearest_sample():
    cap = synth.make_picture(72, 40, chroma_format=1, seed=2)
    ext = pyoracle.extend_border(cap.seq, cap.pre, 32)
    for c, (p, q) in enumerate(zip(cap.pre, ext)):
        xm = ym = 32 if c == 0 else 16
        h, w = p.shape
        assert q.shape == (h + 2 * ym, w + 2 * xm)
        for (y, x) in [(0, 0), (ym - 1, xm + 5), (ym + 3, 0), (ym + h + 2, xm + w + 1), (ym + h - 1, xm + w), (ym + 7, xm + 9)]:
            assert q[y, x] == p[min(max(y - ym, 0), h - 1), min(max(x - xm, 0), w - 1)]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,cf,bd,ctu", [(512, 384, 1, 10, 128), (456, 264, 3, 8, 128), (448, 256, 2, 12, 64), (320, 320, 0, 10, 64), (1920, 1080, 1, 10, 128)])
def test_gpu_lmcs_folded_into_the_first_stage(w, h, cf, bd, ctu):
    from vvc_b200 import gpu
    cap = synth.make_picture(w, h, chroma_format=cf, bit_depth=bd, ctu_size=ctu, seed=w + bd, density=0.8)
    lut = _lut(bd, w)
    # oracle: map the luma plane, then the chain
    mapped = [p.copy() for p in cap.pre]
    pyoracle.lmcs_inverse(mapped, lut)
    saved = cap.pre
    cap.pre = mapped
    want = pyoracle.filter_capture(cap)
    cap.pre = saved
    ctx = gpu.Context(cap.seq, capacity=1)
    try:
        ctx.set_capture(0, cap)                 # uploads the UNMAPPED planes
        ctx.set_lmcs(0, lut)
        ctx.filter(0, 1)
        got = ctx.download(0)
        for c in range(len(got)):
            assert np.array_equal(got[c], want["final"][c]), "fused chain, component %d" % c
        # staged: deblocking alone carries the mapping; SAO and ALF then read mapped samples
        ctx.rewind(0, 1)
        ctx.deblock(0, 1)
        got = ctx.download(0)
        for c in range(len(got)):
            assert np.array_equal(got[c], want["dbf"][c]), "deblocking stage, component %d" % c
        # ALF alone on a reshaped-domain picture is refused, not mis-filtered
        ctx.rewind(0, 1)
        with pytest.raises(gpu.VtmGpuError):
            ctx.alf(0, 1)
        # switching the table off restores the plain chain
        ctx.set_lmcs(0, None)
        ctx.rewind(0, 1)
        ctx.filter(0, 1)
        plain = pyoracle.filter_capture(cap)["final"]
        got = ctx.download(0)
        for c in range(len(got)):
            assert np.array_equal(got[c], plain[c])
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,cf,margin", [(512, 384, 1, 288), (456, 264, 3, 144), (24, 136, 1, 32), (200, 8, 2, 160), (320, 320, 0, 288), (3840, 2160, 1, 288)])
def test_gpu_download_with_border_extension(w, h, cf, margin):
    from vvc_b200 import gpu
    cap = synth.make_picture(w, h, chroma_format=cf, seed=w + margin, density=0.5)
    ctx = gpu.Context(cap.seq, capacity=1)
    try:
        ctx.set_capture(0, cap)
        ctx.filter(0, 1)
        body = ctx.download(0)
        want = pyoracle.extend_border(cap.seq, body, margin)
        got = ctx.download_extended(0, margin)
        for c in range(len(want)):
            assert np.array_equal(got[c], want[c]), "component %d: %d samples differ" % (c, int((got[c] != want[c]).sum()))
    finally:
        ctx.close()
