#!/usr/bin/env python3
"""oracle/make_streams.py -- TEST INFRASTRUCTURE (not product code).

Generates the seeded synthetic 10-bit YUV clips of SURVEY.md Appendix B / BASELINE.md section 4 and
encodes them with the UNMODIFIED reference encoder built by oracle/Makefile
(oracle/_ref/EncoderApp, MD5 decoded-picture-hash SEI on), then decode-verifies each stream with
the unmodified reference decoder (oracle/_ref/DecoderApp must print (OK) for every picture).

Outputs go to oracle/_ref/streams/<name>.{yuv,bin,rec.yuv,enc.log}: git-ignored, but they
travel to the GPU box with the gpurun snapshot.  Needs /root/reference/cfg (encoder cfg files),
so it only runs in the build container.

usage: make_streams.py NAME [NAME...]     (names: see STREAMS below;  'small' = the parity corpus)
"""
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VVC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "streams")
ENC = os.path.join(HERE, "_ref", "EncoderApp")
DEC = os.path.join(HERE, "_ref", "DecoderApp")

# fast encoder flags "K" of SURVEY.md Appendix B (8.5x faster; drops affine/ISP/SBT coverage)
K = ("--MaxMTTHierarchyDepth=1 --MaxMTTHierarchyDepthISliceL=1 --MaxMTTHierarchyDepthISliceC=1 --MTS=0 "
     "--SBT=0 --LFNST=0 --ISP=0 --MIP=0 --MRL=0 --Affine=0 --BIO=0 --DMVR=0 --Geo=0 --CIIP=0 --MMVD=0 "
     "--SMVD=0 --PROF=0 --IMV=0 --BCW=0 --AffineAmvr=0 --SearchRange=32").split()

RA, AI, LD = "encoder_randomaccess_vtm.cfg", "encoder_intra_vtm.cfg", "encoder_lowdelay_vtm.cfg"

# name: (W, H, chroma, frames, seed, sigma, qp, cfgs, extra flags, frame-skip)
STREAMS = {
    # parity corpus (full CTC tool set: affine / ISP / SBT / dual tree / CIIP ... edges exercised)
    "ra_416x240":   (416, 240, 420, 8, 1234, 14, 32, [RA], [], 0),
    "ai_416x240":   (416, 240, 420, 2, 77, 14, 32, [AI], ["--TemporalSubsampleRatio=1"], 0),
    "ld444_416x240": (416, 240, 444, 4, 444, 14, 30, [LD, "444/yuv444.cfg"], [], 0),
    "ra_q22_416x240": (416, 240, 420, 5, 99, 20, 22, [RA], [], 0),
    # chroma correlated with the luma texture (seed >= 9000 selects that content): makes the encoder pick CC-ALF
    "ai_cc_416x240": (416, 240, 420, 2, 9001, 14, 27, [AI], ["--TemporalSubsampleRatio=1"], 0),
    "ld444_cc_416x240": (416, 240, 444, 3, 9002, 14, 27, [LD, "444/yuv444.cfg"], [], 0),
    "ld_q37_832x480": (832, 480, 420, 3, 5, 10, 37, [LD], K, 0),
    # BASELINE.json configs 2..5 (fast flags K)
    "ra_1080p":     (1920, 1080, 420, 32, 4321, 14, 32, [RA], K, 0),
    "ra_2160p_a":   (3840, 2160, 420, 64, 9160, 14, 32, [RA], K, 0),     # frames 0..31
    "ra_2160p_b":   (3840, 2160, 420, 64, 9160, 14, 32, [RA], K, 32),    # frames 32..63
    "ai_4320p":     (7680, 4320, 420, 1, 4320, 14, 32, [AI], K + ["--TemporalSubsampleRatio=1"], 0),
    "ld444_1080p":  (1920, 1080, 444, 16, 444, 14, 32, [LD, "444/yuv444.cfg"], K, 0),
    # picture partitioning with in-loop filtering across the partition boundaries switched off (ALF clip / pad
    # path, SAO availability, deblocking edge suppression): 4x4 tiles in one slice; 3x3 tiles in two raster-scan slices (1 + 8
    # tiles: the second slice wraps around the first one, which exercises the raster-slice corner padding)
    "tiles_832x480": (832, 480, 420, 5, 31, 14, 32, [RA], K + ["--EnablePicPartitioning=1", "--TileColumnWidthArray=2,2", "--TileRowHeightArray=1,1",
                                                               "--DisableLoopFilterAcrossTiles=1"], 0),
    "slices_832x480": (832, 480, 420, 5, 32, 14, 32, [RA], K + ["--EnablePicPartitioning=1", "--TileColumnWidthArray=2,2,3", "--TileRowHeightArray=1,1,2",
                                                                "--RasterScanSlices=1", "--RasterSliceSizes=1,8", "--DisableLoopFilterAcrossSlices=1",
                                                                "--DisableLoopFilterAcrossTiles=0"], 0),
    "slices45_832x480": (832, 480, 420, 3, 33, 14, 30, [RA], K + ["--EnablePicPartitioning=1", "--TileColumnWidthArray=2,2,3", "--TileRowHeightArray=1,1,2",
                                                                  "--RasterScanSlices=1", "--RasterSliceSizes=4,5", "--DisableLoopFilterAcrossSlices=1",
                                                                  "--DisableLoopFilterAcrossTiles=0"], 0),     # slice 0 ends mid tile row: bottom-right corner padding
    # LADF (luma adaptive deblocking QP offset): tc / beta depend on reconstructed samples next to the edge
    "ladf_832x480": (832, 480, 420, 5, 34, 14, 32, [RA], K + ["--LADF=1"], 0),
    # signalled virtual boundaries (360-degree video tool): no in-loop filtering across x = 200 (inside a CTU), x = 512 (a CTU
    # boundary), y = 248 (4 rows above the ALF virtual boundary of its CTU row), y = 384 (a CTU boundary)
    "vb_832x480": (832, 480, 420, 5, 35, 14, 32, [RA], K + ["--LoopFilterAcrossVirtualBoundariesDisabledFlag=1", "--NumVerVirtualBoundaries=2",
                                                            "--NumHorVirtualBoundaries=2", "--VirtualBoundariesPosX=200,512", "--VirtualBoundariesPosY=248,384"], 0),
    # 4:2:2: chroma deblocking grid 16 x 8 luma samples, its own chroma QP mapping, ALF / CC-ALF with sx = 1, sy = 0
    "ld422_416x240": (416, 240, 422, 4, 422, 14, 30, [LD], ["--ChromaFormatIDC=422"], 0),
    # sequence-level variants of the hot path: CTU 64, 12-bit internal depth (this reference snapshot cannot decode its own 8-bit streams), non-zero deblocking offsets
    "ctu64_416x240": (416, 240, 420, 4, 64, 14, 32, [RA], K + ["--CTUSize=64"], 0),
    "ctu32_416x240": (416, 240, 420, 4, 32, 14, 32, [RA], K + ["--CTUSize=32"], 0),
    "bd12_416x240": (416, 240, 420, 3, 12, 14, 32, [RA], K + ["--InternalBitDepth=12"], 0),
    # (4:0:0 makes the encoder of this reference snapshot abort, TypeDef.h:1229: monochrome is covered by seeded pictures vs the oracle only)
    "dbfoffs_416x240": (416, 240, 420, 4, 66, 14, 34, [LD], K + ["--LoopFilterBetaOffset_div2=3", "--LoopFilterTcOffset_div2=-4"], 0),
    # screen-content tools on noise-free blocky 4:4:4 content: palette CUs are not deblocked on their own side (bPartP/QNoFilter),
    # BDPCM edges get bS 0 when both sides use it, IBC CUs take the motion test; P slices (one reference list)
    "scc444_416x240": (416, 240, 444, 3, 8445, 0, 27, [LD, "444/yuv444.cfg"], ["--PLT=1", "--IBC=1", "--BDPCM=1", "--HashME=1"], 0),
    "ldp_416x240": (416, 240, 420, 4, 67, 14, 32, ["encoder_lowdelay_P_vtm.cfg"], K, 0),
    # full CTC tool set at a size with 128-wide CUs (affine / ATMVP sub-block edges inside large CUs, long filters clamped next to affine)
    "ra_full_832x480": (832, 480, 420, 6, 36, 12, 30, [RA], [], 0),
    # LMCS with the slice reshaper ON (the SDR analysis of the encoder leaves it off on the synthetic content above; the PQ signal type always
    # reshapes): executeLoopFilters maps the luma reconstruction through the inverse table before the deblocking (DecLib.cpp:570-577)
    "lmcs_416x240": (416, 240, 420, 6, 68, 14, 32, [RA], K + ["--LMCSSignalType=1"], 0),
    "lmcs_full_832x480": (832, 480, 420, 4, 69, 12, 30, [RA], ["--LMCSSignalType=1"], 0),
    # short 4K clip for fast turnaround (first 8 pictures of the config-3 source)
    "ra_2160p_8":   (3840, 2160, 420, 8, 9160, 14, 32, [RA], K, 0),
}


def gen_yuv(path, W, H, chroma, frames, seed, sigma):
    """Seeded synthetic clip, planar little-endian u16, values 0..1023."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    sx = 2 if chroma in (420, 422) else 1
    sy = 2 if chroma == 420 else 1
    cy, cx = np.mgrid[0:H // sy, 0:W // sx]
    cs = 1.0 if chroma == 420 else 2.0   # keep chroma feature sizes similar at full resolution (4:2:2: stretched vertically, fine)
    with open(path, "wb") as f:
        for t in range(frames):
            y = (512 + 260 * np.sin((xx + 6 * t) / 41.0) * np.cos((yy - 4 * t) / 29.0)
                 + 140 * (((xx + 9 * t) // 48 + (yy // 40)) % 2)
                 + 60 * np.sin((xx * yy) / 9000.0 + t) + rng.normal(0, sigma, (H, W)))
            u = (512 + 200 * np.sin((cx / cs - 3 * t) / 31.0) * np.cos(cy / cs / 57.0)
                 + rng.normal(0, sigma / 2, cy.shape))
            v = (512 + 200 * np.cos((cy / cs + 2 * t) / 23.0) + 60 * (((cx / cs + 4 * t) // 24) % 2)
                 + rng.normal(0, sigma / 2, cy.shape))
            if 8000 <= seed < 9000:   # screen content: a few flat colours in sharp-edged rectangles and thin lines (palette / IBC friendly)
                y = 128.0 + 256 * (((xx + 3 * t) // 24 + yy // 18) % 3) + 300 * ((xx % 16 == 5) | (yy % 20 == 7))
                u = 300.0 + 200 * (((cx + 2 * t) // 20 + cy // 14) % 3)
                v = 700.0 - 180 * ((cx // 28 + (cy + t) // 10) % 4)
            if seed >= 9000:   # CC-ALF content: chroma carries a blurred copy of the luma structure
                yd = y[::sy, ::sx]
                u = 512 + 0.45 * (yd - 512) + rng.normal(0, sigma / 2, cy.shape)
                v = 512 - 0.35 * (yd - 512) + 40 * np.sin(cx / 17.0) + rng.normal(0, sigma / 2, cy.shape)
            for p in (y, u, v):
                f.write(np.clip(p, 0, 1023).astype("<u2").tobytes())


def make(name):
    W, H, chroma, frames, seed, sigma, qp, cfgs, extra, skip = STREAMS[name]
    os.makedirs(OUT, exist_ok=True)
    base = os.path.join(OUT, name)
    yuv = os.path.join(OUT, f"src_{W}x{H}_{chroma}_s{seed}_f{frames}.yuv")
    if not os.path.exists(yuv):
        gen_yuv(yuv + ".tmp", W, H, chroma, frames, seed, sigma)
        os.replace(yuv + ".tmp", yuv)
    nenc = frames - skip if skip else (frames if name != "ra_2160p_a" else 32)
    cmd = [ENC]
    for c in cfgs:
        cmd += ["-c", os.path.join(REF, "cfg", c)]
    cmd += ["-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(nenc), "-fs", str(skip),
            "--InputBitDepth=10", "--InternalBitDepth=10", f"--InputChromaFormat={chroma}",
            "-q", str(qp), "--SEIDecodedPictureHash=1", "-b", base + ".bin", "-o", base + ".rec.yuv"] + extra
    t0 = time.time()
    with open(base + ".enc.log", "w") as log:
        log.write(" ".join(cmd) + "\n")
        log.flush()
        subprocess.run(cmd, stdout=log, stderr=subprocess.STDOUT, check=True)
    t1 = time.time()
    r = subprocess.run([DEC, "-b", base + ".bin", "-o", base + ".dec.yuv", "-d", "0"],
                       capture_output=True, text=True)
    ok = r.stdout.count("(OK)")
    bad = r.stdout.count("ERROR")
    same = open(base + ".dec.yuv", "rb").read() == open(base + ".rec.yuv", "rb").read()
    with open(base + ".dec.log", "w") as log:
        log.write(r.stdout + r.stderr)
    os.remove(base + ".dec.yuv")
    print(f"{name}: encode {t1 - t0:.0f}s, decoder rc={r.returncode} MD5 OK={ok} ERR={bad} "
          f"recon==decode:{same}", flush=True)
    if r.returncode != 0 or bad or not same:
        raise SystemExit(f"{name}: reference decoder rejected the stream")


if __name__ == "__main__":
    names = sys.argv[1:]
    if names == ["small"]:
        names = ["ra_416x240", "ai_416x240", "ld444_416x240"]
    for n in names:
        make(n)
