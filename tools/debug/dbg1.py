import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
from vvc_b200 import gpu, synth
which = sys.argv[1]
cf = 0 if which == "luma" else 1
cap = synth.make_picture(416, 240, chroma_format=cf, seed=7, density=0.9)
ctx = gpu.Context(cap.seq)
ctx.set_capture(0, cap)
if which == "sao": ctx.sao(0,1)
else: ctx.deblock(0,1)
print(which, "ok")
