"""CPU: every picture of every small parity stream (hand-picked and random encoder configurations) through the reference backend
(oracle/_ref/DecoderApp_cap, the reference's own filter classes, capture mode) and through the shim's derivation + the plain-C oracle:
the planes after deblocking, after SAO and after ALF must be identical.  tools/check_oracle_all.py does the same for the large
streams."""
import glob
import os
import subprocess
import sys

import pytest

from conftest import GOLDEN
from vvc_b200 import capture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
DEC = os.path.join(ROOT, "oracle", "_ref", "DecoderApp_cap")
SMALL = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "streams", "*.bin")) if os.path.getsize(p) < 64 * 1024)


@pytest.mark.skipif(not os.path.exists(DEC), reason="oracle/_ref/DecoderApp_cap not built (needs the reference sources)")
@pytest.mark.parametrize("stream", SMALL)
def test_oracle_equals_reference_on_every_picture(stream, tmp_path):
    import check_oracle_all
    env = dict(os.environ, VTMGPU_SHIM_BACKEND="ref", VTMGPU_CAPTURE_DIR=str(tmp_path))
    r = subprocess.run([DEC, "-b", os.path.join(GOLDEN, "streams", stream), "-d", "0"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ERROR" not in r.stdout and r.stdout.count("(OK)") > 0, r.stdout[-1000:] + r.stderr[-1000:]
    caps = sorted(n for n in os.listdir(tmp_path) if n.endswith(".cap"))
    assert len(caps) == r.stdout.count("(OK)")
    for n in caps:
        bad = check_oracle_all.check(capture.load(os.path.join(str(tmp_path), n)))
        assert not bad, "%s %s: oracle differs from the reference at %s" % (stream, n, bad)
