import glob
import hashlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


_cache = {}


def load_golden(name):
    from vvc_b200 import capture
    if name not in _cache:
        _cache[name] = capture.load(os.path.join(GOLDEN, name))
    return _cache[name]


def plane_md5(plane):
    """Decoded-picture-hash MD5 of one plane as the reference computes it for > 8-bit content
    (PicYuvMD5.cpp:45-87,188-212): samples as little-endian 16-bit, raster order."""
    return hashlib.md5(plane.astype("<u2").tobytes()).hexdigest()
