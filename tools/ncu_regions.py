#!/usr/bin/env python3
"""tools/ncu_regions.py -- per-device-function and per-opcode view of an ncu report, using the INLINE chains of nvdisasm.

tools/ncu_functions.py attributes the SASS of CUDA-header intrinsics to "the preceding function"; this tool reads
`nvdisasm --print-line-info-inline`, takes for every SASS instruction the innermost frame that lies in one of the repo's own
sources and names the device function enclosing that line.  Prints executed warp instructions (and thread instructions per luma
pixel when PIXELS is given), stall samples and the opcode mix per function.

usage: ncu_regions.py REPORT.ncu-rep CUBIN NCU_KERNEL_REGEX:CUBIN_SECTION_SUBSTRING SRC_DIR [PIXELS]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

FUNC = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:__global__|__device__|static|inline|__host__|__forceinline__|__noinline__|\s)+[\w:<>\*&\s]+?\b(\w+)\s*\(")


def functions_of(path):
    """line number -> enclosing function name (first definition line at column 0 above it)"""
    names, cur = {}, "(file scope)"
    try:
        src = open(path).read().splitlines()
    except OSError:
        return names
    pending_template = False
    for i, ln in enumerate(src, 1):
        if ln.startswith("template"):
            pending_template = True
        if ln.startswith(("__device__", "__global__", "template", "static __device__", "inline __device__")):
            m = FUNC.match(ln)
            if m:
                cur = m.group(1)
        names[i] = cur
    return names


def main():
    rep, cubin, kern, srcdir = sys.argv[1:5]
    pixels = float(sys.argv[5]) if len(sys.argv) > 5 else None
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
            break
        if len(r) == len(hdr):
            data.append(r)
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout
    funcs = {}
    frames, insts, on, chain = [], [], False, []
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            on = sec in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            chain.append((m.group(1), int(m.group(2))))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            if chain:
                frames = chain
            chain = []
            insts.append(list(frames))
    if len(insts) != len(data):
        print("warning: %d SASS instructions in the cubin section, %d in the report" % (len(insts), len(data)), file=sys.stderr)
    n = min(len(insts), len(data))
    per_f = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for k in range(n):
        r = data[k]
        ni = int(r[ix["Instructions Executed"]])
        ns = int(r[ix["# Samples"]] or 0)
        name = "(unknown)"
        for f, l in insts[k]:
            if os.path.abspath(f).startswith(os.path.abspath(srcdir)) or os.path.basename(f) in os.listdir(srcdir):
                p = os.path.join(srcdir, os.path.basename(f))
                if p not in funcs:
                    funcs[p] = functions_of(p)
                name = funcs[p].get(l, "?")
                break
        s = r[ix["Source"]].split()
        op = s[1] if s[0].startswith("@") else s[0]
        op = op.rstrip(";")
        e = per_f[name]
        e[0] += ni; e[1] += ns; e[2][op] += ni
        tot_i += ni; tot_s += ns
    print("kernel %s: %d executed warp instructions, %d samples" % (kern, tot_i, tot_s))
    for name, (ni, ns, ops) in sorted(per_f.items(), key=lambda kv: -kv[1][0]):
        if ni == 0:
            continue
        extra = "  %6.2f thr-inst/px" % (ni * 32 / pixels) if pixels else ""
        print("%5.1f%% inst %5.1f%% smp%s  %s" % (100.0 * ni / tot_i, 100.0 * ns / max(tot_s, 1), extra, name))
        print("        " + ", ".join("%s %.1f%%" % (o, 100.0 * c / ni) for o, c in ops.most_common(8)))


if __name__ == "__main__":
    main()
