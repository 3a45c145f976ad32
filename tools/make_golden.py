#!/usr/bin/env python3
"""tools/make_golden.py -- generates the committed golden fixtures under tests/golden/ (run in the build container).

For each small reference-encoded stream (oracle/make_streams.py) the stream is decoded by oracle/_ref/DecoderApp_cap
= the reference decoder + our host shim with VTMGPU_SHIM_BACKEND=ref, i.e. every filter stage is executed by the
REFERENCE's own classes (RefLoopFilter / RefSampleAdaptiveOffset / RefAdaptiveLoopFilter, unmodified sources) while the
shim records what crosses the drop-in boundary: pre-filter planes, the flattened side information, and the planes
after each stage.  Selected pictures are stored compressed (.npz, stage planes as deltas) together with the
decoded-picture MD5s the reference ENCODER put into the SEI (printed "(OK)" by the decoder), in manifest.json.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from vvc_b200 import capture  # noqa: E402

DEC = os.path.join(ROOT, "oracle", "_ref", "DecoderApp_cap")
STREAMS = os.path.join(ROOT, "oracle", "_ref", "streams")
OUT = os.path.join(ROOT, "tests", "golden")

# stream -> decode-order indices of the pictures to keep
SELECT = {
    "ra_416x240": [0, 1, 7],           # I picture + B pictures of different temporal layers (affine/SBT/CIIP edges)
    "ai_416x240": [1],                 # dual-tree intra, SAO + ALF active
    "ai_cc_416x240": [0],              # chroma correlated with luma: CC-ALF on for both chroma components
    "ld444_416x240": [2],              # 4:4:4, chroma SAO/ALF at full resolution
    "ld444_cc_416x240": [1],           # 4:4:4 with CC-ALF
    "ra_q22_416x240": [0, 2],          # low QP: many APS filter sets, SAO on
    "ld_q37_832x480": [1],             # high QP low delay (strong / long deblocking filters)
    "tiles_832x480": [0, 2],           # 4x4 tiles, no in-loop filtering across tile boundaries (ALF clip path, SAO availability)
    "slices45_832x480": [0],           # slices of 4 + 5 tiles: bottom-right corner padding
    "ladf_832x480": [0, 2],            # LADF: luma records carry QPs, thresholds derived from the samples
    "vb_832x480": [0, 2],              # signalled virtual boundaries inside CTUs and on CTU edges
    "ld422_416x240": [0, 2],           # 4:2:2: chroma deblocking grid / QP mapping, ALF and CC-ALF with sx = 1, sy = 0
    "ctu64_416x240": [0, 2],           # CTU 64: tile = CTU in k_alf, 16 chroma rows per CTU row in the SAO / ALF boundary logic
    "ctu32_416x240": [0, 2],           # CTU 32: four CTUs per 64x64 ALF tile
    "bd12_416x240": [0, 1],            # 12-bit: tc / beta scaling, clip tables, IDP operand ranges
    "dbfoffs_416x240": [1, 3],         # slice-level beta / tc offsets
    "scc444_416x240": [0, 1],          # palette / IBC / BDPCM on blocky 4:4:4 content
    "ldp_416x240": [1, 3],             # P slices
    "ra_full_832x480": [1, 4],         # full CTC tool set with 128-wide CUs (affine / sub-block edges, long filters)
    "slices_832x480": [0, 3],          # 3x3 tiles in two raster-scan slices, no filtering across slices (ALF corner padding)
}


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    mpath = os.path.join(OUT, "manifest.json")
    only = sys.argv[1:]                       # optional: regenerate only these streams, keep the other manifest entries
    if only and os.path.exists(mpath):
        with open(mpath) as f:
            manifest = {k: v for k, v in json.load(f).items() if v["stream"] not in only}
    for stream, keep in SELECT.items():
        if only and stream not in only:
            continue
        bs = os.path.join(STREAMS, stream + ".bin")
        if not os.path.exists(bs):
            print("skip (no stream):", stream)
            continue
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, VTMGPU_SHIM_BACKEND="ref", VTMGPU_CAPTURE_DIR=tmp)
            r = subprocess.run([DEC, "-b", bs, "-d", "0"], env=env, capture_output=True, text=True)
            assert r.returncode == 0 and "ERROR" not in r.stdout, r.stdout + r.stderr
            md5 = {}
            for m in re.finditer(r"POC\s+(\d+).*?\[MD5:([0-9a-f]+),([0-9a-f]+),([0-9a-f]+),\(OK\)\]", r.stdout):
                md5[int(m.group(1))] = [m.group(2), m.group(3), m.group(4)]
            caps = sorted(n for n in os.listdir(tmp) if n.endswith(".cap"))
            for i in keep:
                cap = capture.load(os.path.join(tmp, caps[i]))
                name = "%s_%s.npz" % (stream, caps[i][:-4])
                cap.save_npz(os.path.join(OUT, name))
                manifest[name] = dict(stream=stream, decode_index=i, poc=cap.seq["poc"], sei_md5=md5[cap.seq["poc"]],
                                      seq=cap.seq, activity=cap.activity())
                print(name, os.path.getsize(os.path.join(OUT, name)) // 1024, "KiB", manifest[name]["activity"])
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
