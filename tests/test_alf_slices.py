"""Pictures whose slices carry different ALF parameters (ALFProcess reloads the APS data at every slice change,
AdaptiveLoopFilter.cpp:429-441).  The reference encoder writes the same parameters into every slice of a picture, so this path is
pinned against the oracle on seeded multi-slice descriptions: CPU tests for the oracle's composition, GPU tests for
vtmgpu_set_alf_slices through the C ABI (k_alf, k_alf<false> and k_alf_parts)."""
import numpy as np
import pytest

import pyoracle
from vvc_b200 import synth


def _sao_output(cap):
    return pyoracle.filter_capture(cap, stages=("dbf", "sao"))["final"]


def test_oracle_equal_slices_reduce_to_one_set():
    cap = synth.make_picture(256, 256, chroma_format=1, ctu_size=64, seed=5, density=0.9)
    want = pyoracle.filter_capture(cap)["final"]
    p = cap.alf_params()
    ctu_slice = (np.arange(cap.num_ctus) * 3 // cap.num_ctus).astype(np.uint8)
    got = _sao_output(cap)
    pyoracle.alf_slices(cap.seq, got, [p, p, p], ctu_slice)
    for c in range(3):
        assert np.array_equal(got[c], want[c])


def test_oracle_slices_are_filtered_with_their_own_sets():
    cap = synth.make_picture(256, 256, chroma_format=1, ctu_size=64, seed=6, density=1.0)
    single = pyoracle.filter_capture(cap)["final"]
    slices, ctu_slice = synth.make_alf_slices(cap, 3, seed=1)
    got = _sao_output(cap)
    pyoracle.alf_slices(cap.seq, got, slices, ctu_slice)
    wc = cap.width_in_ctus
    differs = False
    for a in range(cap.num_ctus):
        y0, x0 = (a // wc) * 64, (a % wc) * 64
        same = np.array_equal(got[0][y0:y0 + 64, x0:x0 + 64], single[0][y0:y0 + 64, x0:x0 + 64])
        if ctu_slice[a] == 0:
            assert same, "CTU %d of slice 0 must equal the single-set result" % a
        differs |= not same
    assert differs


GPU_CASES = [(512, 384, 1, 128, 3, None), (456, 264, 3, 128, 2, 1), (448, 256, 2, 64, 4, None), (416, 240, 1, 32, 3, 0), (640, 384, 1, 128, 6, None),
             (1920, 1080, 1, 128, 5, 2)]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,cf,ctu,ns,off", GPU_CASES)
def test_gpu_alf_slices_vs_oracle(w, h, cf, ctu, ns, off):
    from vvc_b200 import gpu
    cap = synth.make_picture(w, h, chroma_format=cf, ctu_size=ctu, seed=w + ns, density=0.9, partitions=(ctu == 64))
    slices, ctu_slice = synth.make_alf_slices(cap, ns, seed=w, off_slice=off)
    want = _sao_output(cap)
    pyoracle.alf_slices(cap.seq, want, slices, ctu_slice)
    ctx = gpu.Context(cap.seq, capacity=1)
    try:
        ctx.set_capture(0, cap)
        ctx.set_alf_slices(0, slices, ctu_slice)
        ctx.filter(0, 1)
        got = ctx.download(0)
        for c in range(len(want)):
            assert np.array_equal(got[c], want[c]), "component %d differs (%d samples)" % (c, int((got[c] != want[c]).sum()))
        # the single-set entry point afterwards replaces the slice tables again
        ctx.upload(0, cap.pre)
        ctx.set_alf(0, cap.alf_params())
        ctx.filter(0, 1)
        one = pyoracle.filter_capture(cap)["final"]
        got = ctx.download(0)
        for c in range(len(one)):
            assert np.array_equal(got[c], one[c])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_alf_slices_rejects_too_many_sets():
    from vvc_b200 import gpu
    cap = synth.make_picture(1024, 256, chroma_format=1, ctu_size=64, seed=2, density=0.9)
    slices, ctu_slice = synth.make_alf_slices(cap, 10, seed=3)
    ctx = gpu.Context(cap.seq, capacity=1)
    try:
        with pytest.raises(gpu.VtmGpuError):
            ctx.set_alf_slices(0, slices, ctu_slice)
    finally:
        ctx.close()
