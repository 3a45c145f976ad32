#!/usr/bin/env python3
"""tools/bench_bands.py -- BASELINE config 4: ONE 7680x4320 intra picture filtered by N GPUs in CTU-row bands with the halo
exchange over NCCL/NVLink (vvc_b200/bands.py).  Launch: python -m torch.distributed.run --nproc-per-node N tools/bench_bands.py

Device-resident replay: every rank keeps its band of the captured picture in HBM (rewind), one iteration =
deblocking+SAO kernel -> halo exchange (4 rows per plane and border) -> ALF kernel; timed with barrier + synchronize on
both sides, max over ranks.  Prints one JSON line (Mpixel/s of the whole picture) and checks the assembled output against
the MD5 the reference encoder put into the stream when --verify is given."""
import argparse
import glob
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--stream", default=os.path.join(ROOT, "tests", "golden", "streams", "ai_4320p.bin"))
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--graph", action="store_true", help="record one iteration (kernels, halo copies, NCCL send/recv) in a CUDA graph and replay it")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from vvc_b200 import abi, bands, capture, gpu
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp(prefix="vtmgpu_cap8k_")
    try:
        env = dict(os.environ, VTMGPU_SHIM_BACKEND="gpu", VTMGPU_CAPTURE_DIR=tmp, VTMGPU_DEVICE=str(local))
        r = subprocess.run([os.path.join(ROOT, "vvc_b200", "_bin", "DecoderApp_gpu"), "-b", args.stream, "-d", "0"], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "(OK)" in r.stdout and "ERROR" not in r.stdout, (r.stdout + r.stderr)[-500:]
        cap = capture.load(sorted(glob.glob(os.path.join(tmp, "*.cap")))[0])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    ctx = gpu.Context(cap.seq, capacity=1, device=local)
    out, (y0, y1) = bands.filter_picture_in_bands(cap, ctx, rank, world, dist)       # also uploads band + side info
    ok = None
    if args.verify:
        want = cap.stage["alf"] if cap.stage.get("alf") is not None else None
        sy = abi.chroma_shifts(cap.seq["chroma_format"])[1]
        ok = bool(want is not None and all(np.array_equal(out[c][y0 >> (sy if c else 0):y1 >> (sy if c else 0)], want[c][y0 >> (sy if c else 0):y1 >> (sy if c else 0)])
                                           for c in range(cap.ncomp)))
    bnds = bands.band_rows(cap.height, world)
    sx, sy = abi.chroma_shifts(cap.seq["chroma_format"])
    shifts = [0] + ([sy, sy] if cap.ncomp == 3 else [])
    widths = [cap.width] + [cap.width >> sx] * (cap.ncomp - 1)
    dev = torch.device("cuda", local)
    plan = bands.halo_plan_packed(bnds, rank, shifts)
    bufs = [torch.empty(sum(widths) * bands.HALO, dtype=torch.int16, device=dev) for _ in plan]
    # everything of one iteration on torch's current stream: kernels, halo copies and the NCCL send/recv order on the device
    ctx.set_stream(torch.cuda.current_stream().cuda_stream, True)

    def step():
        ctx.rewind(0, 1)
        ctx.deblock_sao(0, 1)
        bands.exchange_packed(plan, ctx, dist, bufs, host_sync=False)
        ctx.alf(0, 1)

    run, graph_ok = step, None
    if args.graph and world > 1:
        raise SystemExit("--graph: with this torch / NCCL build (2.11 / 2.28.9) the run hangs when the halo send/recv is part of the capture; single GPU only")
    if args.graph:
        # the whole iteration as ONE graph launch: the host cost of an iteration (Python, two kernel launches, four halo copies, one
        # NCCL group) is what bounds the band mode, not the 245 KB exchange.  Capture on a side stream that the context is pointed at.
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ctx.set_stream(side.cuda_stream, True)
            for _ in range(3):
                step()                                   # NCCL sets up its point-to-point channels outside the capture
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                step()
            g.replay()
            side.synchronize()
            got = [np.zeros_like(p) for p in out]
            ctx.download_rows(0, got, y0, y1)
            side.synchronize()
        run = g.replay
        sy_ = abi.chroma_shifts(cap.seq["chroma_format"])[1]
        graph_ok = all(np.array_equal(got[c][y0 >> (sy_ if c else 0):y1 >> (sy_ if c else 0)], out[c][y0 >> (sy_ if c else 0):y1 >> (sy_ if c else 0)]) for c in range(cap.ncomp))
    for _ in range(args.warmup):
        run()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    oks = torch.tensor([1 if ok in (None, True) and graph_ok in (None, True) else 0], device=dev)
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    if rank == 0:
        px = cap.width * cap.height
        print(json.dumps({"metric": "DBF+SAO+ALF Mpixel/s, one 7680x4320 picture in CTU-row bands", "value": round(px / float(t.item()) / 1e6, 1), "unit": "Mpixel/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(float(t.item()) * 1e3, 4), "scaling": "strong",
                          "bands": bnds, "halo_rows": bands.HALO, "halo_bytes_per_border": sum(widths) * bands.HALO * 2 * 2,
                          "verified_vs_single_gpu_capture": None if ok is None else bool(oks.item()),
                          "cuda_graph": bool(args.graph), "graph_replay_equals_direct": None if graph_ok is None else bool(oks.item()), "activity": cap.activity()}))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
