#!/usr/bin/env python3
"""tools/sass_summary.py -- opcode histogram of every kernel in libvtmgpu.so (cuobjdump -sass; no GPU needed).

usage: sass_summary.py [LIB.so] > profiles/rN_sass_summary.txt
Shows per kernel: SASS instruction count, registers / shared memory from the ELF, the architecture, the TMA / bulk-copy / mbarrier
mnemonics that prove asynchronous tile movement (UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier), the packed
16-bit integer instructions the filters run on, and the 25 most frequent opcodes."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vvc_b200", "csrc", "libvtmgpu.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", ln)
    if m and cur:
        usage[cur] = m.groups()
kern, arch = None, None
hist = collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("library:", os.path.relpath(lib), " arch:", arch)
KEY = ["UTMALDG", "UBLKCP", "SYNCS", "VIADDMNMX", "VIADD.16x2", "VIMNMX", "IDP", "PRMT", "LDS", "STS", "LDG", "STG", "BAR", "ATOM", "RED", "HMMA", "UTCMMA"]
for k, h in hist.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    tot = sum(h.values())
    u = usage.get(k)
    print("\n== %s  [%d SASS instructions%s]" % (name, tot, ", %s registers, %s B stack, %s B static smem" % u if u else ""))
    keys = []
    for key in KEY:
        n = sum(c for op, c in h.items() if op.startswith(key))
        if n:
            keys.append("%s %d" % (key, n))
    print("   key mnemonics: " + ", ".join(keys))
    print("   top opcodes  : " + ", ".join("%s %d" % (op, c) for op, c in h.most_common(25)))
