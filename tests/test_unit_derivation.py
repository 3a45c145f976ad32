"""SURVEY 8f n1, CPU part: the deblocking derivation from the flattened block structure (include/vtmgpu_derive.h -- the source the kernel
k_dbf_derive runs) against the CU walk that restates LoopFilter::xDeblockCU (vvc_b200/shim/vtm_flatten.cpp, pinned to the reference's
own filters by the golden fixtures and the decoder MD5s).  The test build of the decoder flattens every picture of every small parity
stream, runs the unit derivation on the host and compares all four record arrays with the walk's: byte-identical, no picture may need
the walk instead.  The GPU test tests/test_gpu_parity.py::test_decoder_device_derivation does the same with the kernel's output."""
import glob
import os
import re
import subprocess

import pytest

from conftest import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEC = os.path.join(ROOT, "oracle", "_ref", "DecoderApp_cap")
SMALL = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "streams", "*.bin")) if os.path.getsize(p) < 256 * 1024)


@pytest.mark.skipif(not os.path.exists(DEC), reason="oracle/_ref/DecoderApp_cap not built (needs the reference sources)")
@pytest.mark.parametrize("stream", SMALL)
def test_unit_derivation_equals_cu_walk(stream):
    env = dict(os.environ, VTMGPU_SHIM_BACKEND="ref", VTMGPU_SHIM_CHECK_UNITS="1", VTMGPU_SHIM_TIMING="1")
    r = subprocess.run([DEC, "-b", os.path.join(GOLDEN, "streams", stream), "-d", "0"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ERROR" not in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    ok = r.stdout.count("(OK)")
    m = re.search(r"units_checked=(\d+)", r.stdout)
    assert ok > 0 and m and int(m.group(1)) == ok, r.stdout[-500:]
