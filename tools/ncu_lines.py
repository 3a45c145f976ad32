#!/usr/bin/env python3
"""tools/ncu_lines.py -- per-source-line view of an ncu report (no GPU needed).

Joins `ncu -i REP --page source --csv` (per-SASS-instruction counters, in program order) with the line annotations of
`nvdisasm --print-line-info CUBIN` for the same kernel and prints the hottest CUDA source lines: share of executed
warp instructions, share of stall samples and the dominant stall reason.

usage: ncu_lines.py REPORT.ncu-rep CUBIN KERNEL_REGEX[:CUBIN_SECTION_SUBSTRING] [TOP_N]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sec = kern.split(":")[1] if ":" in kern else kern       # substring of the (mangled) cubin section, e.g. k_alfILb0 for k_alf<false>
    kern = kern.split(":")[0]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {n: i for i, n in enumerate(hdr)}
    data = []
    for r in rows[h + 1:]:
        if r and r[0] == "Address":
            break                                           # a second result of the same kernel
        if len(r) == len(hdr):
            data.append(r)
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]

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
    if len(lines) != len(data):
        print("warning: %d SASS instructions in the cubin, %d in the report -- cubin does not match" % (len(lines), len(data)))
    n = min(len(lines), len(data))
    agg = defaultdict(lambda: [0, 0, defaultdict(int), 0])
    tot_i = tot_s = 0
    for k in range(n):
        r = data[k]
        ie, sm = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        a = agg[lines[k]]
        a[0] += ie
        a[1] += sm
        a[3] += 1
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                a[2][c[6:]] += v
        tot_i += ie
        tot_s += sm
    print("kernel %s: %d SASS instructions, %d executed warp instructions, %d samples" % (kern, n, tot_i, tot_s))
    src_cache = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
        text = ""
        if key:
            if key[0] not in src_cache:
                try:
                    src_cache[key[0]] = open("/root/repo/vvc_b200/csrc/" + key[0]).read().splitlines()
                except OSError:
                    src_cache[key[0]] = []
            s = src_cache[key[0]]
            text = s[key[1] - 1].strip()[:90] if key[1] - 1 < len(s) else ""
        print("%5.1f%% smp %5.1f%% inst %4d sass  %-28s %-22s %s" % (100.0 * a[1] / max(tot_s, 1), 100.0 * a[0] / max(tot_i, 1), a[3],
              "%s:%d" % key if key else "?", ",".join("%s:%d" % (k, v) for k, v in st), text))


if __name__ == "__main__":
    main()
