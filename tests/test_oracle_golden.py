"""CPU: pins the plain-C oracle (oracle/vvc_filters_oracle.c) to the REFERENCE.

The fixtures in tests/golden/ were produced by tools/make_golden.py: pre-filter planes and flattened side info
captured at the drop-in boundary, and the planes after each stage as computed by the reference's own filter classes
(unmodified LoopFilter.cpp / SampleAdaptiveOffset.cpp / AdaptiveLoopFilter.cpp), plus the decoded-picture MD5 the
reference ENCODER wrote into the SEI.  Bar: bit-exact.
"""
import numpy as np
import pytest

import pyoracle
from conftest import golden_names, load_golden, plane_md5

NAMES = golden_names()


def test_golden_present():
    assert len(NAMES) >= 8


@pytest.mark.parametrize("name", NAMES)
def test_oracle_stage_by_stage(name):
    cap = load_golden(name)
    # each stage starts from the REFERENCE's previous-stage output, so an error cannot hide behind another stage
    planes = [p.copy() for p in cap.pre]
    pyoracle.deblock(cap.seq, planes, cap.deblock_params())
    for c in range(cap.ncomp):
        assert np.array_equal(planes[c], cap.stage["dbf"][c]), "deblocking differs in component %d" % c
    prev = cap.stage["dbf"]
    if cap.stage["sao"] is not None:
        planes = [p.copy() for p in prev]
        ctus = cap.sao_ctus()
        pyoracle.sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, *cap.sao_scale)
        pyoracle.sao(cap.seq, planes, ctus, cap.vb_struct())
        for c in range(cap.ncomp):
            assert np.array_equal(planes[c], cap.stage["sao"][c]), "SAO differs in component %d" % c
        prev = cap.stage["sao"]
    if cap.stage["alf"] is not None:
        planes = [p.copy() for p in prev]
        pyoracle.alf(cap.seq, planes, cap.alf_params())
        for c in range(cap.ncomp):
            assert np.array_equal(planes[c], cap.stage["alf"][c]), "ALF differs in component %d" % c


@pytest.mark.parametrize("name", NAMES)
def test_oracle_chain_matches_sei_md5(name, manifest):
    cap = load_golden(name)
    final = pyoracle.filter_capture(cap)["final"]
    assert [plane_md5(p) for p in final] == manifest[name]["sei_md5"]


def test_golden_covers_every_tool(manifest):
    """The fixture set must exercise every tool of the path at least once."""
    acts = [m["activity"] for m in manifest.values()]
    assert any(a.get("sao_on_frac_c0", 0) > 0 for a in acts) and any(a.get("sao_on_frac_c1", 0) > 0 for a in acts)
    assert any(a.get("alf_on_frac_c0", 0) > 0 for a in acts) and any(a.get("alf_on_frac_c1", 0) > 0 for a in acts)
    assert any(a.get("ccalf_on_frac_c1", 0) > 0 or a.get("ccalf_on_frac_c2", 0) > 0 for a in acts)
    assert any(m["seq"]["chroma_format"] == 3 for m in manifest.values())
    kinds = set()
    for n in NAMES:
        cap = load_golden(n)
        for d in range(2):
            r = cap.dbf_luma[d]
            r = r[(r & 0x7FF) != 0]
            kinds |= set(((r >> 22) & 7).tolist()) | set(((r >> 25) & 7).tolist())
    assert {1, 3, 7} <= kinds, "filter lengths seen: %s" % sorted(kinds)
