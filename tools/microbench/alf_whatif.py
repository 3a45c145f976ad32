#!/usr/bin/env python3
"""tools/microbench/alf_whatif.py -- where k_alf's TIME goes (run on the GPU box).

Times the chain on two seeded 4K pictures (16 slots) with libraries built with EXTRA=-DALF_WHATIF=n (alf_kernel.cuh: parts of the
kernel left out -- the results are wrong, only the time matters): one process per library (VTMGPU_LIB).
usage: alf_whatif.py LIB.so [LIB.so ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    from vvc_b200 import gpu, synth
    cache = "/tmp/alf_whatif_caps.npz"
    caps = [synth.make_picture(3840, 2160, seed=2160 + i, density=0.6) for i in range(2)]
    n = 16
    ctx = gpu.Context(caps[0].seq, capacity=n, device=0)
    for s in range(n):
        ctx.set_capture(s, caps[s % 2])
    ctx.sync()
    ctx.set_profiling(True)
    for _ in range(3):
        ctx.rewind(0, n)
        ctx.filter(0, n, sync=False)
    ctx.sync()
    tot = [0.0, 0.0]
    for _ in range(10):
        ctx.rewind(0, n)
        ctx.filter(0, n, sync=False)
        ms = ctx.stage_ms()
        tot[0] += ms[0] / 10
        tot[1] += ms[1] / 10
    print("%-40s k_dbf_sao %.1f us  k_alf %.1f us per picture" % (os.path.basename(os.environ.get("VTMGPU_LIB", "default")), tot[0] / n * 1e3, tot[1] / n * 1e3), flush=True)
    ctx.close()


if __name__ == "__main__":
    if os.environ.get("WHATIF_CHILD"):
        child()
    else:
        for lib in sys.argv[1:]:
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, WHATIF_CHILD="1", VTMGPU_LIB=os.path.abspath(lib)))
