#!/usr/bin/env python3
"""tools/ncu_functions.py -- per-device-function view of an ncu report (no GPU needed).

Like tools/ncu_lines.py, but buckets the executed warp instructions and the stall samples of a kernel by the device function
(of SOURCE.cuh) the SASS instruction was inlined from; instructions of CUDA headers are attributed to the preceding function
("(intrinsics)"), kernel-body lines to blocks of 20 source lines.

usage: ncu_functions.py REPORT.ncu-rep CUBIN NCU_KERNEL_REGEX[:CUBIN_SECTION_SUBSTRING] SOURCE.cuh [TOP_N]
"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, cubin, kern, srcf = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    sec = kern.split(":")[1] if ":" in kern else kern
    kern = kern.split(":")[0]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {n: i for i, n in enumerate(hdr)}
    data = []
    for r in rows[h + 1:]:
        if r and r[0] == "Address":
            break                                   # a second result of the same kernel
        if len(r) == len(hdr):
            data.append(r)
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    lines, cur, on = [], None, False
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            on = sec in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    here = os.path.dirname(os.path.abspath(__file__))
    path = srcf if os.path.exists(srcf) else os.path.join(here, "..", "vvc_b200", "csrc", srcf)
    funcs = []
    for i, l in enumerate(open(path).read().splitlines()):
        m = re.match(r"(?:template <[^>]*> )?__(?:device|global)__ .*?(\w+)\(", l)
        if m:
            funcs.append((i + 1, m.group(1)))
        if l.startswith("__global__"):
            funcs.append((i + 1, "KERNEL"))

    def fn(line):
        name = "?"
        for s, n in funcs:
            if s <= line:
                name = n
        return name

    base = os.path.basename(srcf)
    agg, smp, tot, ts, last = defaultdict(int), defaultdict(int), 0, 0, "?"
    for k in range(min(len(data), len(lines))):
        ie, s = int(data[k][ix["Instructions Executed"]]), int(data[k][ix["# Samples"]])
        f, l = lines[k] if lines[k] else ("?", 0)
        if f == base:
            key = fn(l)
            if key in ("KERNEL", kern):
                key = "KERNEL:%d" % (l // 20 * 20)
            last = key
        else:
            key = last + " (intrinsics)"
        agg[key] += ie
        smp[key] += s
        tot += ie
        ts += s
    print("kernel %s: %d executed warp instructions, %d samples, %d SASS instructions" % (kern, tot, ts, len(lines)))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s" % (100 * v / tot, 100 * smp[k] / max(ts, 1), k))


if __name__ == "__main__":
    main()
