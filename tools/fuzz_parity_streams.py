#!/usr/bin/env python3
"""tools/fuzz_parity_streams.py -- TEST TOOL (build container): random small reference-encoded streams for the parity sweep.

Draws N random encoder configurations (RA / LD / AI, QP 22..42, 4:2:0 / 4:2:2 / 4:4:4, three random tool switches out of: dual
tree, ISP / SBT / MTS, affine / SbTMVP, IBC, palette, CIIP / GEO / MMVD, joint CbCr, BDPCM, LMCS, SAO / ALF / CC-ALF / deblocking
off, CTU 32 / 64, 12-bit, LADF, deblocking offsets, tiles, raster-scan slices, signalled virtual boundaries), encodes the seeded
synthetic clip with the unmodified reference encoder (oracle/make_streams.py) and decode-verifies it with the reference decoder.
Then run  tools/check_oracle_all.py oracle/_ref/streams/fuzz<SEED>_*.bin  : the shim's host derivation + the oracle must
reproduce the reference's planes after every stage of every picture.

usage: fuzz_parity_streams.py SEED [N]"""
import sys, os, itertools, subprocess, random
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import make_streams as ms
from concurrent.futures import ThreadPoolExecutor
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
RA, AI, LD = ms.RA, ms.AI, ms.LD
opts = [
 ("--DualITree=0",), ("--ISP=1", "--SBT=1", "--MTS=1"), ("--Affine=1", "--SbTMVP=1", "--PROF=1"), ("--IBC=1",), ("--CIIP=1", "--Geo=1", "--MMVD=1"),
 ("--JointCbCr=1",), ("--BDPCM=1",), ("--LMCSEnable=0",), ("--SAO=0",), ("--ALF=0",), ("--CCALF=0",), ("--LoopFilterDisable=1",),
 ("--DepQuant=0",), ("--MaxMTTHierarchyDepth=3",), ("--LFNST=1", "--MIP=1"), ("--TemporalSubsampleRatio=1",),
 ("--CTUSize=64",), ("--CTUSize=32",), ("--LADF=1",), ("--InternalBitDepth=12",), ("--LoopFilterBetaOffset_div2=-2", "--LoopFilterTcOffset_div2=3"),
 ("--EnablePicPartitioning=1", "--TileColumnWidthArray=1,2", "--TileRowHeightArray=1", "--DisableLoopFilterAcrossTiles=1"),
 ("--EnablePicPartitioning=1", "--TileColumnWidthArray=2", "--TileRowHeightArray=1", "--RasterScanSlices=1", "--RasterSliceSizes=1,3", "--DisableLoopFilterAcrossSlices=1"),
 ("--LoopFilterAcrossVirtualBoundariesDisabledFlag=1", "--NumVerVirtualBoundaries=1", "--NumHorVirtualBoundaries=1", "--VirtualBoundariesPosX=136", "--VirtualBoundariesPosY=72"),
 ("--LoopFilterAcrossVirtualBoundariesDisabledFlag=1", "--NumVerVirtualBoundaries=2", "--NumHorVirtualBoundaries=0", "--VirtualBoundariesPosX=64,256"),
 ("--PLT=1", "--IBC=1"),
 ("--DMVR=1", "--BIO=1"), ("--SMVD=1", "--BCW=1", "--IMV=1", "--AffineAmvr=1", "--Affine=1"), ("--MRL=1",), ("--TransformSkip=1", "--ChromaTS=1", "--BDPCM=1"),
 ("--LMChroma=0",), ("--WeightedPredP=1", "--WeightedPredB=1"), ("--LMCSSignalType=1",), ("--LMCSSignalType=1", "--DualITree=1"),
]
names = []
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 10):
    cfg = random.choice([RA, LD, AI])
    qp = random.choice([22, 27, 32, 37, 42])
    chroma = random.choice([420, 420, 420, 444, 422])
    extra = list(ms.K)
    for o in random.sample(opts, 3):
        for kv in o:
            key = kv.split("=")[0]
            extra = [e for e in extra if not e.startswith(key + "=")]
            extra.append(kv)
    cfgs = [cfg] + (["444/yuv444.cfg"] if chroma == 444 else [])
    if chroma == 422: extra.append("--ChromaFormatIDC=422")
    w, h = random.choice([(416, 240), (352, 288), (264, 200), (480, 272)])
    name = "fuzz%s_%d" % (sys.argv[1] if len(sys.argv) > 1 else "1", i)
    frames = 2 if cfg == AI else 4
    if cfg == AI and "--TemporalSubsampleRatio=1" not in extra: extra.append("--TemporalSubsampleRatio=1")
    ms.STREAMS[name] = (w, h, chroma, frames, 7000 + i + 100 * int(sys.argv[1] if len(sys.argv) > 1 else 1), random.choice([0, 8, 14, 20]), qp, cfgs, extra, 0)
    names.append(name)
def run(n):
    try:
        ms.make(n); return n, "ok"
    except BaseException as e:
        return n, "FAILED " + str(e)[:100]
with ThreadPoolExecutor(8) as ex:
    for n, r in ex.map(run, names):
        print(n, r, ms.STREAMS[n][:3], ms.STREAMS[n][6], [e for e in ms.STREAMS[n][8] if e not in ms.K], flush=True)
