#!/usr/bin/env python3
"""tools/microbench/e2e_variants.py -- what bounds vtmgpu_batch_filter (run on the GPU box).

Needs a library built with EXTRA=-DVTMGPU_BATCH_KNOBS (an experiment build: VTMGPU_BATCH_SKIP leaves steps of the per-picture
sequence out; bits: 1 record lists, 8 the kernels, 16 download, 32 plane upload).  Prints ms per 64-picture step and variant.
Round 2 on B200: all 37.3 (42.3 before the downloads were aligned with the uploads), no kernels 36.7, no download 32.8, no upload 30.1,
kernels alone 6.9."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from vvc_b200 import abi, gpu
    caps, _ = bench.load_pictures(int(os.environ.get("DISTINCT", "8")), 0)
    B = 64
    seq = caps[0].seq
    pin_in = [[torch.from_numpy(p.copy()).pin_memory() for p in c.pre] for c in caps]
    side = []
    for c in caps:
        ctus = c.sao_ctus()
        if ctus is not None:
            gpu.sao_reconstruct(ctus, c.width_in_ctus, c.ncomp, c.sao_scale[0], c.sao_scale[1])
        dp = gpu.sparse_records(c.dbf_luma, c.dbf_chroma if c.ncomp > 1 else None, pin=True)
        side.append((dp, ctus, c.alf_params(), c.vb_struct()))
    for skip, lanes in [(0, 8), (8, 8), (0, 4), (0, 3), (0, 2), (0, 16), (0, 8)]:
        os.environ["VTMGPU_BATCH_SKIP"] = str(skip)
        batch = gpu.Batch(seq, lanes=lanes, device=0)
        pin_out = [[torch.empty_like(t).pin_memory() for t in pin_in[0]] for _ in range(lanes)]
        pics = [gpu.host_picture([t.numpy() for t in pin_in[k % len(caps)]], [t.numpy() for t in pin_out[k % lanes]], side[k % len(caps)][0],
                                 side[k % len(caps)][1], side[k % len(caps)][2], side[k % len(caps)][3]) for k in range(B)]
        arr = (abi.HostPicture * B)(*pics)
        batch.filter(arr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 4
        for _ in range(n):
            batch.filter(arr)
        print("skip=%-3d lanes=%-2d  %.2f ms per step" % (skip, lanes, (time.perf_counter() - t0) / n * 1e3), flush=True)
        batch.close()


if __name__ == "__main__":
    main()
