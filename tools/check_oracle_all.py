#!/usr/bin/env python3
"""tools/check_oracle_all.py -- TEST TOOL (build container, no GPU): the plain-C oracle against the reference on EVERY picture of
the small parity streams, not only on the committed fixtures.  Each stream is decoded by oracle/_ref/DecoderApp_cap with the
reference's own filter classes as backend and the shim in capture mode; for every captured picture the oracle must reproduce
the reference's planes after deblocking, after SAO and after ALF (each stage fed with the reference's previous stage).

usage: check_oracle_all.py [STREAM.bin ...]      (default: all 416x240 / 832x480 streams under tests/golden/streams)"""
import glob
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402
from vvc_b200 import capture  # noqa: E402

DEC = os.path.join(ROOT, "oracle", "_ref", "DecoderApp_cap")


def check(cap):
    planes = [p.copy() for p in cap.pre]
    pyoracle.deblock(cap.seq, planes, cap.deblock_params())
    bad = [("dbf", c) for c in range(cap.ncomp) if not np.array_equal(planes[c], cap.stage["dbf"][c])]
    prev = cap.stage["dbf"]
    if cap.stage["sao"] is not None:
        planes = [p.copy() for p in prev]
        ctus = cap.sao_ctus()
        pyoracle.sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, *cap.sao_scale)
        pyoracle.sao(cap.seq, planes, ctus, cap.vb_struct())
        bad += [("sao", c) for c in range(cap.ncomp) if not np.array_equal(planes[c], cap.stage["sao"][c])]
        prev = cap.stage["sao"]
    if cap.stage["alf"] is not None:
        planes = [p.copy() for p in prev]
        pyoracle.alf(cap.seq, planes, cap.alf_params())
        bad += [("alf", c) for c in range(cap.ncomp) if not np.array_equal(planes[c], cap.stage["alf"][c])]
    return bad


def main():
    streams = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "streams", "*_416x240.bin")) +
                                     glob.glob(os.path.join(ROOT, "tests", "golden", "streams", "*_832x480.bin")))
    total = failed = 0
    for bs in streams:
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, VTMGPU_SHIM_BACKEND="ref", VTMGPU_CAPTURE_DIR=tmp)
            r = subprocess.run([DEC, "-b", bs, "-d", "0"], env=env, capture_output=True, text=True)
            assert r.returncode == 0 and "ERROR" not in r.stdout, r.stdout[-500:] + r.stderr[-500:]
            names = sorted(n for n in os.listdir(tmp) if n.endswith(".cap"))
            res = []
            for n in names:
                bad = check(capture.load(os.path.join(tmp, n)))
                total += 1
                failed += bool(bad)
                res.append("ok" if not bad else str(bad))
            print("%-28s %2d pictures: %s" % (os.path.basename(bs), len(names), "all ok" if all(x == "ok" for x in res) else res), flush=True)
    print("%d pictures, %d differ" % (total, failed))
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
