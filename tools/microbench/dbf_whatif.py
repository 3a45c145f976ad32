#!/usr/bin/env python3
"""tools/microbench/dbf_whatif.py -- k_dbf_sao's time by stage on the VTM-encoded 4K pictures (run on the GPU box):
deblocking + SAO, deblocking alone, SAO alone (= the kernel's skeleton: tiles in by TMA, strips out, no passes, no pass barriers)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from vvc_b200 import gpu

caps, _ = bench.load_pictures(8, 0)
n = 32
ctx = gpu.Context(caps[0].seq, capacity=n, device=0)
for s in range(n):
    ctx.set_capture(s, caps[s % len(caps)])
ctx.sync()
for name, call in (("deblocking + SAO", ctx.deblock_sao), ("deblocking", ctx.deblock), ("SAO", ctx.sao)):
    for _ in range(3):
        ctx.rewind(0, n)
        call(0, n)
    ctx.sync()
    tot = 0.0
    for _ in range(10):
        ctx.rewind(0, n)
        ctx.sync()
        ctx.timer_start()
        call(0, n)
        tot += ctx.timer_stop()
    print("%-18s %.1f us per picture" % (name, tot / 10 / n * 1e3), flush=True)
ctx.close()
