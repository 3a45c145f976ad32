#!/usr/bin/env python3
"""tools/ncu_summary.py -- one CSV row per ncu --set full report (no GPU needed) + profiles/traffic.json for bench.py.

usage: ncu_summary.py PICTURES_IN_LAUNCH OUT.csv TRAFFIC.json KERNEL=REPORT.ncu-rep [KERNEL=REPORT.ncu-rep ...]
Columns: duration, DRAM bytes read / written (per launch) and per luma pixel, DRAM throughput % of peak, executed warp instructions and
thread instructions per luma pixel, issue-slot utilisation, ALU / FMA / LSU pipe utilisation, shared-memory bank conflicts / wavefronts,
barrier stall per issue, occupancy, registers, shared memory."""
import csv
import io
import json
import subprocess
import sys

pics = int(sys.argv[1])
out_csv, out_json = sys.argv[2], sys.argv[3]
PX = 3840 * 2160
WANT = [("gpu__time_duration.sum", "duration_us"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"), ("smsp__inst_executed.sum", "warp_instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier_stall_per_issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "short_scoreboard_stall_per_issue"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "math_pipe_throttle_per_issue"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"), ("launch__registers_per_thread", "registers"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem_per_block"), ("launch__grid_size", "grid"), ("launch__occupancy_limit_registers", "ctas_per_sm_by_registers"),
        ("launch__occupancy_limit_shared_mem", "ctas_per_sm_by_smem")]
rows, traffic = [], {}
for arg in sys.argv[4:]:
    kern, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = r[0], r[1], r[2]
    ix = {h: i for i, h in enumerate(hdr)}
    row = {"kernel": kern, "pictures_in_launch": pics, "report": rep}
    for metric, name in WANT:
        if metric in ix:
            v = vals[ix[metric]].replace(",", "")
            try:
                v = float(v)
                u = units[ix[metric]]
                if name in ("dram_read", "dram_write"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                if name == "duration_us":
                    v *= {"ns": 1e-3, "us": 1, "usecond": 1, "msecond": 1e3, "ms": 1e3, "second": 1e6}.get(u, 1)
            except ValueError:
                pass
            row[name] = v
    px = PX * pics
    row["dram_bytes_per_luma_pixel"] = round((row["dram_read"] + row["dram_write"]) / px, 3)
    row["thread_instructions_per_luma_pixel"] = round(row["warp_instructions"] * 32 / px, 2)
    row["smem_conflict_frac"] = round(row["smem_bank_conflicts"] / max(row["smem_wavefronts"], 1), 3)
    row["us_per_picture"] = round(row["duration_us"] / pics, 2)
    rows.append(row)
    traffic[kern] = {"dram_bytes_read": row["dram_read"], "dram_bytes_write": row["dram_write"], "pictures_in_launch": pics,
                     "dram_bytes_per_luma_pixel": row["dram_bytes_per_luma_pixel"], "duration_us": row["duration_us"], "issue_active_pct": row.get("issue_active_pct"),
                     "warp_instructions": row["warp_instructions"], "report": rep,
                     "source": "ncu --set full --clock-control none of ONE %d-picture launch of the bench loop (tools/profile_round.sh)" % pics}
keys = list(rows[0].keys())
with open(out_csv, "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=keys)
    w.writeheader()
    for row in rows:
        w.writerow(row)
json.dump(traffic, open(out_json, "w"), indent=1)
for row in rows:
    print({k: row[k] for k in ("kernel", "us_per_picture", "dram_bytes_per_luma_pixel", "thread_instructions_per_luma_pixel", "issue_active_pct", "smem_conflict_frac", "barrier_stall_per_issue", "pipe_alu_pct", "pipe_fma_pct")})
