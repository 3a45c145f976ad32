"""CPU: the host side of the drop-in boundary (vvc_b200/shim: deblocking derivation, SAO / ALF flattening, partition and
virtual-boundary flags, LADF records, record lists) against the committed fixtures.

oracle/_ref/DecoderApp_cap = the reference decoder + our shim with the REFERENCE's own filter classes as backend
(VTMGPU_SHIM_BACKEND=ref, no GPU needed).  In capture mode the shim writes everything that crosses the C ABI; the side
information of the fixture pictures must be byte-identical to tests/golden/*.npz, every picture must pass the decoder's MD5
check, and the shim's self-check "record lists == dense record arrays" must hold (it THROWs otherwise)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from vvc_b200 import capture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEC = os.path.join(ROOT, "oracle", "_ref", "DecoderApp_cap")
STREAMS = ["ra_416x240", "ra_full_832x480", "ld422_416x240", "ctu64_416x240", "ctu32_416x240", "bd12_416x240", "dbfoffs_416x240", "scc444_416x240", "ldp_416x240", "tiles_832x480", "slices_832x480", "slices45_832x480", "ladf_832x480", "vb_832x480"]
PLANE_PREFIXES = ("pre_", "dbf_", "sao_", "alf_")


def _side_sections(cap):
    return {k: bytes(v) for k, v in cap._sections.items() if not (k[:4] in PLANE_PREFIXES and k[4:].isdigit())}


@pytest.mark.skipif(not os.path.exists(DEC), reason="oracle/_ref/DecoderApp_cap not built (needs the reference sources)")
@pytest.mark.parametrize("stream", STREAMS)
def test_shim_side_info_matches_fixtures(stream, manifest, tmp_path):
    bs = os.path.join(GOLDEN, "streams", stream + ".bin")
    env = dict(os.environ, VTMGPU_SHIM_BACKEND="ref", VTMGPU_CAPTURE_DIR=str(tmp_path))
    r = subprocess.run([DEC, "-b", bs, "-d", "0"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ERROR" not in r.stdout and r.stdout.count("(OK)") > 0
    caps = sorted(n for n in os.listdir(tmp_path) if n.endswith(".cap"))
    assert r.stdout.count("(OK)") == len(caps)
    checked = 0
    for name, m in manifest.items():
        if m["stream"] != stream:
            continue
        got = capture.load(os.path.join(str(tmp_path), caps[m["decode_index"]]))
        want = load_golden(name)
        a, b = _side_sections(got), _side_sections(want)
        assert sorted(a) == sorted(b), (sorted(a), sorted(b))
        for k in a:
            assert a[k] == b[k], "%s: section %s differs from the fixture" % (name, k)
        for c in range(got.ncomp):
            assert np.array_equal(got.pre[c], want.pre[c])
        checked += 1
    assert checked > 0
