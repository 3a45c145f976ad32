#!/usr/bin/env python3
"""tools/ncu_phases.py -- where the executed instructions and stall samples of a kernel go (no GPU needed).

Joins `ncu -i REP --page source --csv` with `nvdisasm --print-line-info-inline CUBIN`.  Every SASS instruction is
attributed to (a) the device function it was inlined from (innermost non-CUDA-header frame) and (b) the opcode class;
prints thread-instructions per luma pixel when PIXELS is given.

usage: ncu_phases.py REPORT.ncu-rep CUBIN NCU_KERNEL_REGEX:CUBIN_SECTION_SUBSTRING [PIXELS] [--lines FILE_SUBSTRING]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def func_table(path):
    """line -> name of the enclosing top-level function of a .cuh (heuristic: a line that starts a definition at column 0)."""
    names, cur = {}, "?"
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|inline|__host__).*?([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    try:
        src = open(path).read().splitlines()
    except OSError:
        return names
    for i, ln in enumerate(src, 1):
        m = pat.match(ln)
        if m and not ln.startswith(" "):
            cur = m.group(1)
            g = re.search(r"\b(k_[A-Za-z0-9_]+)\s*\(", ln)
            if "__global__" in ln and g:
                cur = g.group(1)
        names[i] = cur
    return names


def main():
    rep, cubin, kern = sys.argv[1:4]
    rest = sys.argv[4:]
    pixels = float(rest[0]) if rest and not rest[0].startswith("--") else None
    want_lines = rest[rest.index("--lines") + 1] if "--lines" in rest else None
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
    insts, frames, on = [], [], False
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            on = sec in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if "inlined at" not in ln and frames and frames[-1][2]:
                pass
            frames.append((m.group(1), int(m.group(2)), "inlined at" in ln))
            if "inlined at" not in ln:
                cur_frames = frames
                frames = []
                last = cur_frames
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            insts.append((m.group(2).strip(), list(last) if 'last' in dir() else []))
    if len(insts) != len(data):
        sys.stderr.write("warning: %d SASS instructions in the cubin section, %d in the report\n" % (len(insts), len(data)))
    tables = {}
    by_fn, by_op, by_line = defaultdict(lambda: [0, 0]), defaultdict(int), defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for (text, fr), r in zip(insts, data):
        n = int(float(r[ix["Thread Instructions Executed"]] or 0))
        smp = int(r[ix["# Samples"]] or 0)
        tot_i += n
        tot_s += smp
        fn = "?"
        inner = None
        for f, l, _ in fr:                       # innermost first
            if "/usr/local/cuda" in f or "targets/" in f or f.endswith("packed16.cuh") or f.endswith("async_copy.cuh"):
                continue
            inner = (f, l)
            break
        if inner:
            if inner[0] not in tables:
                tables[inner[0]] = func_table(inner[0])
            fn = tables[inner[0]].get(inner[1], "?")
            if fn.startswith("k_") and fr:
                outer = fr[-1]
                fn = "%s:%d" % (fn, outer[1] // 10 * 10)
            if want_lines and want_lines in inner[0]:
                by_line[(inner[0].split("/")[-1], inner[1])][0] += n
                by_line[(inner[0].split("/")[-1], inner[1])][1] += smp
        op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]
        by_fn[fn][0] += n
        by_fn[fn][1] += smp
        by_op[op] += n
    scale = (1.0 / pixels) if pixels else 100.0 / max(tot_i, 1)
    unit = "inst/px" if pixels else "% inst"
    print("kernel %s: %.0f thread instructions%s, %d samples" % (kern, tot_i, (" = %.1f per pixel" % (tot_i / pixels)) if pixels else "", tot_s))
    for fn, (n, smp) in sorted(by_fn.items(), key=lambda kv: -kv[1][0]):
        if n * 200 < tot_i and smp * 200 < tot_s:
            continue
        print("  %7.2f %s  %5.1f%% smp  %s" % (n * scale, unit, 100.0 * smp / max(tot_s, 1), fn))
    print("opcodes:", ", ".join("%s %.1f" % (k, v * scale) for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:24]))
    if want_lines:
        for (f, l), (n, smp) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:40]:
            print("  %7.2f %s  %5.1f%% smp  %s:%d" % (n * scale, unit, 100.0 * smp / max(tot_s, 1), f, l))


if __name__ == "__main__":
    main()
