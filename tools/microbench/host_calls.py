"""Host-side cost of the per-picture C ABI calls at 3840x2160 (what the end-to-end loop of bench.py pays per picture)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from vvc_b200 import gpu, synth
cap = synth.make_picture(3840, 2160, seed=1, density=0.9)
ctx = gpu.Context(cap.seq, capacity=1)
pin = [torch.from_numpy(p.copy()).pin_memory() for p in cap.pre]
out = [torch.empty_like(t).pin_memory() for t in pin]
ctus = cap.sao_ctus(); gpu.sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, 0, 0)
dp, ap = cap.deblock_params(), cap.alf_params()
def t(name, fn, n=30):
    fn(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
    print("%-28s issue %.3f ms   incl. device %.3f ms" % (name, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
t("upload (async, pinned)", lambda: ctx.upload(0, [x.numpy() for x in pin], sync=False))
t("set_deblock (staged)", lambda: ctx.set_deblock(0, dp))
t("set_deblock_async", lambda: ctx.set_deblock(0, dp, sync=False))
t("set_sao", lambda: ctx.set_sao(0, ctus))
t("set_alf", lambda: ctx.set_alf(0, ap))
t("filter (async)", lambda: ctx.filter(0, 1, sync=False))
t("download (async, pinned)", lambda: ctx.download(0, [x.numpy() for x in out], sync=False))
sp = gpu.sparse_records(cap.dbf_luma, cap.dbf_chroma, pin=True)
t("set_deblock_sparse", lambda: ctx.set_deblock_sparse(0, sp))
