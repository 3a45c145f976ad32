#!/usr/bin/env python3
"""bench.py -- DBF+SAO+ALF(+CC-ALF) throughput of the VVC in-loop filter chain at 3840x2160 10-bit 4:2:0.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                    (the reference's own CPU filters on the host cores)

One step = one pass of the whole chain over a batch of independent pictures that are already resident in HBM
(picture-parallel: every rank filters its own batch, no data-path collective => weak scaling).  Prints ONE JSON line.

  value      Mpixel/s (luma pixels through DBF+SAO+ALF), device-resident, CUDA-event timed on the launching stream,
             barrier + synchronize on both sides, max over ranks
  e2e        the same metric through the C ABI with HOST buffers: per picture pinned H2D of the planes + side-info
             pack/upload + the chain + D2H of the filtered planes, all inside the timed region
  roofline   dominant kernel: algorithmic HBM bytes per launch / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the reference's filter classes (oracle/_ref/DecoderApp_cap, reference backend) timed on the host cores

Workload data: VTM-encoded synthetic 4K pictures captured at the drop-in boundary when data/captures/ra_2160p exists,
otherwise seeded synthetic pictures with synthetic side info (vvc_b200/synth.py); stated in config.workload.
"""
import argparse
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
METRIC = "DBF+SAO+ALF Mpixel/s at 3840x2160 10-bit"
WORKLOAD = "3840x2160 10-bit 4:2:0 random-access (QP32, CC-ALF) in-loop filter chain DBF+SAO+ALF+CC-ALF, picture-parallel"
B_ALG_CHAIN = 6.5          # algorithmic bytes per luma pixel of the whole chain at 4:2:0 (SURVEY.md 8d / BASELINE.md 3)
B_ALG_DBF = 6.5            # k_dbf_sao : read 3 + write 3 + 0.5 segment records
B_ALG_SAOALF = 6.0         # k_alf     : read 3 + write 3 (CTU parameters negligible)
STREAM_4K = os.path.join(ROOT, "tests", "golden", "streams", "ra_2160p_8.bin")
DEC_GPU = os.path.join(ROOT, "vvc_b200", "_bin", "DecoderApp_gpu")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------------------
def load_pictures(distinct, device):
    """Returns (list of Capture, description).

    Preferred workload (BASELINE config 3): the VTM-encoded 3840x2160 RA stream of synthetic YUV (tests/golden/streams) is
    decoded HERE by the reference decoder with our filters dropped in (vvc_b200/_bin/DecoderApp_gpu, every picture must
    print MD5 (OK)); the shim's capture mode records what crosses the boundary per picture -- pre-filter planes and the
    flattened side information -- and those pictures are replayed.  Fallback: seeded synthetic pictures."""
    import shutil
    import tempfile
    from vvc_b200 import capture, synth
    if os.path.exists(STREAM_4K) and os.path.exists(DEC_GPU) and not os.environ.get("VTMGPU_BENCH_SYNTH"):
        tmp = tempfile.mkdtemp(prefix="vtmgpu_cap_")
        try:
            env = dict(os.environ, VTMGPU_SHIM_BACKEND="gpu", VTMGPU_CAPTURE_DIR=tmp, VTMGPU_DEVICE=str(device))
            r = subprocess.run([DEC_GPU, "-b", STREAM_4K, "-d", "0"], env=env, capture_output=True, text=True, timeout=600)
            ok = r.stdout.count("(OK)")
            files = sorted(glob.glob(os.path.join(tmp, "*.cap")))
            if r.returncode == 0 and ok == len(files) and ok >= 1 and "ERROR" not in r.stdout:
                caps = [capture.load(f) for f in files[:distinct]]
                return caps, ("captured: VTM-encoded (RA, QP32, CC-ALF) synthetic 4K YUV decoded by DecoderApp_gpu, %d/%d pictures MD5 (OK), "
                              "%d distinct pictures replayed" % (ok, len(files), len(caps)))
            sys.stderr.write("bench.py: DecoderApp_gpu capture failed (rc=%d, OK=%d): %s\n" % (r.returncode, ok, (r.stdout + r.stderr)[-400:]))
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    caps = [synth.make_picture(W4K, H4K, seed=2160 + i, density=0.6) for i in range(distinct)]
    return caps, "synthetic planes + synthetic side info (vvc_b200/synth.py density 0.6), %d distinct pictures" % distinct


def activity_summary(caps):
    keys = sorted({k for c in caps for k in c.activity() if "frac" in k})
    acts = [c.activity() for c in caps]
    return {k: round(sum(a.get(k, 0.0) for a in acts) / len(acts), 3) for k in keys}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own filter classes on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(max_procs=None, repeats=1):
    """One unmodified-reference decoder process per host core (oracle/_ref/DecoderApp_cap with the reference backend: the
    three filter stages are executed by RefLoopFilter / RefSampleAdaptiveOffset / RefAdaptiveLoopFilter, AVX2 ALF),
    all started together on the 4K stream; the shim's steady_clock timers bracket ONLY those three reference calls.
    Returns (Mpixel/s aggregate, cores, kind, sample description) or None."""
    dec = os.path.join(ROOT, "oracle", "_ref", "DecoderApp_cap")
    streams = [os.path.join(ROOT, "tests", "golden", "streams", n) for n in ("ra_2160p_8.bin", "ra_1080p.bin", "ra_416x240.bin")]
    stream = next((s for s in streams if os.path.exists(s)), None)
    cores = len(os.sched_getaffinity(0))
    if max_procs:
        cores = min(cores, max_procs)
    if os.path.exists(dec) and stream and os.access(dec, os.X_OK):
        env = dict(os.environ, VTMGPU_SHIM_BACKEND="ref", VTMGPU_SHIM_TIMING="1")
        best = None
        for _ in range(repeats):
            procs = [subprocess.Popen(["taskset", "-c", str(sorted(os.sched_getaffinity(0))[i]), dec, "-b", stream, "-d", "0"], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for i in range(cores)]
            outs = [p.communicate()[0] for p in procs]
            secs, px = [], 0
            for o in outs:
                m = re.search(r"vtmgpu-shim-timing: pictures=(\d+) luma_pixels=(\d+) filter_s=([0-9.eE+-]+)", o)
                if not m or "ERROR" in o:
                    secs = None
                    break
                secs.append(float(m.group(3)))
                px += int(m.group(2))
            if secs:
                v = px / max(secs) / 1e6
                best = v if best is None else max(best, v)
        if best is not None:
            return best, cores, "reference", "%s decoded by %d pinned processes (one per core), filter stages only" % (os.path.basename(stream), cores)
    # fall back to the plain-C port of the oracle (single thread)
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        from vvc_b200 import synth
        cap = synth.make_picture(W4K, H4K, seed=2160, density=0.6)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 8.0:
            pyoracle.filter_capture(cap)
            n += 1
        return n * cap.luma_pixels() / (time.perf_counter() - t0) / 1e6, 1, "port", "%d synthetic 4K pictures through oracle/vvc_filters_oracle.c, 1 thread" % n
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    vals = []
    res = None
    for _ in range(max(1, min(args.steps, 3))):
        res = cpu_reference_run()
        if res is None:
            break
        vals.append(res[0])
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/DecoderApp_cap + data/streams nor the oracle port could run"}))
        return 0
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": len(vals),
            "warmup": 0, "ms_per_step": round((time.perf_counter() - t0) * 1e3 / len(vals), 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arm": "the reference's own filter classes on the host CPU: " + res[3]},
            "cpu_baseline": {"value": round(v, 2), "unit": "Mpixel/s", "cores": res[1], "kind": res[2], "sample": res[3]},
            "e2e": {"value": round(v, 2), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Multi-rank runs: keep this rank's threads -- and with them its page-locked staging buffers (first touch) -- on the NUMA node
    the GPU hangs off; the end-to-end leg moves ~50 MB per picture through host memory.  Returns the node or None."""
    try:
        r = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True, timeout=30)
        bus = r.stdout.strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                   # 00000000:1b:00.0 -> 0000:1b:00.0
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read())
        if node < 0:
            return None
        cpus = set()
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="pictures per step per GPU (BASELINE config 3: 64 frames)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct pictures replicated to fill the batch")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the forced-on synthetic leg (N=1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    from vvc_b200 import gpu

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the filter chain has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 and os.environ.get("VTMGPU_BENCH_NUMA", "1") == "1" else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    caps, workload = load_pictures(args.distinct, local)
    B = args.batch
    seq = caps[0].seq
    px_per_pic = caps[0].luma_pixels()
    ctx = gpu.Context(seq, capacity=B, device=local)
    for s in range(B):
        ctx.set_capture(s, caps[s % len(caps)])
    ctx.sync()
    ctx.set_profiling(True)

    def step():
        ctx.rewind(0, B)
        ctx.filter(0, B, sync=False)

    for _ in range(args.warmup):
        step()
    ctx.sync()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = ctx.launch_count()
    stage = [0.0, 0.0]
    ctx.timer_start()
    for _ in range(args.steps):
        step()
        # stage events are recorded inside filter(); reading them back needs the step to be finished, so the per-kernel
        # figures come from a second, identical timed loop below -- this loop only carries the start/stop events
    ms_total = ctx.timer_stop()
    barrier()
    launches = ctx.launch_count() - launches0
    for _ in range(args.steps):            # per-kernel durations (CUDA events around each launch, same stream)
        step()
        ms = ctx.stage_ms()
        stage[0] += ms[0]
        stage[1] += ms[1]
    clocks = sampler.summary()
    ms_step = max_over_ranks(ms_total / args.steps)
    value = world * B * px_per_pic / (ms_step * 1e-3) / 1e6

    peak, peak_src = measured_peak()
    k_ms = [s / args.steps for s in stage]
    dom = 1 if k_ms[1] >= k_ms[0] else 0
    alg_bytes = (B_ALG_SAOALF if dom else B_ALG_DBF) * px_per_pic * B
    achieved = alg_bytes / (k_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_alf" if dom else "k_dbf_sao", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "kernel_ms": {"k_dbf_sao": round(k_ms[0], 4), "k_alf": round(k_ms[1], 4)},
                "chain_achieved": round(B_ALG_CHAIN * px_per_pic * B / (ms_total / args.steps * 1e-3) / 1e9, 1),
                "chain_frac": round(B_ALG_CHAIN * px_per_pic * B / (ms_total / args.steps * 1e-3) / 1e9 / peak, 4)}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            per_px = json.load(open(tr)).get(roofline["kernel"], {}).get("dram_bytes_per_luma_pixel")
            if per_px is not None:
                roofline["traffic"] = round(per_px * px_per_pic * B)      # ncu dram__bytes_read+write per luma pixel x pixels of one launch
        except Exception:
            pass

    # ---- end to end through the C ABI with host buffers -------------------------------------------------------------
    NCTX, CH = int(os.environ.get("VTMGPU_E2E_NCTX", "8")), int(os.environ.get("VTMGPU_E2E_CH", "1"))   # contexts (streams) x slots: copies of one chunk overlap kernels of another
    ectx = [gpu.Context(seq, capacity=CH, device=local) for _ in range(NCTX)]
    pin_in = [[torch.from_numpy(p.copy()).pin_memory() for p in c.pre] for c in caps]
    pin_out = [[[torch.empty_like(t).pin_memory() for t in pin_in[0]] for _ in range(CH)] for _ in range(NCTX)]
    import ctypes as C
    from vvc_b200 import abi
    dense_records = os.environ.get("VTMGPU_E2E_DENSE_RECORDS", "0") == "1"
    side = []
    for c in caps:
        ctus = c.sao_ctus()
        if ctus is not None:
            gpu.sao_reconstruct(ctus, c.width_in_ctus, c.ncomp, c.sao_scale[0], c.sao_scale[1])
        # the decoder-side producer of the segment records (the shim's CU walk) appends the active units to lists in
        # page-locked memory: no staging copy in the library, ~0.15 instead of 0.75 B of records per luma pixel on the bus
        if dense_records:
            dp, keep, nbytes = abi.DeblockParams(), [], 0
            for d in range(2):
                tl = torch.from_numpy(c.dbf_luma[d].view(np.int32).copy()).pin_memory()
                keep.append(tl)
                nbytes += c.dbf_luma[d].nbytes
                dp.luma[d] = C.cast(tl.data_ptr(), C.POINTER(C.c_uint32))
                if c.dbf_chroma[d].size:
                    tc = torch.from_numpy(c.dbf_chroma[d].view(np.int64).copy()).pin_memory()
                    keep.append(tc)
                    nbytes += c.dbf_chroma[d].nbytes
                    dp.chroma[d] = C.cast(tc.data_ptr(), C.POINTER(C.c_uint64))
            dp._keep = keep
        else:
            dp = gpu.sparse_records(c.dbf_luma, c.dbf_chroma if c.ncomp > 1 else None, pin=True)
            nbytes = sum(dp.luma_count[d] * C.sizeof(abi.DbfLumaEntry) + dp.chroma_count[d] * C.sizeof(abi.DbfChromaEntry) for d in range(2))
        side.append((dp, ctus, c.alf_params(), nbytes))
    plane_bytes = sum(t.numel() * 2 for t in pin_in[0])
    h2d = sum(plane_bytes + side[k % len(caps)][3] for k in range(B)) / B       # per picture, averaged over the batch
    d2h = plane_bytes

    issue = [0.0]
    ncalls = [0]

    def e2e_step():
        t_issue = time.perf_counter()
        for base in range(0, B, CH):
            cx = ectx[(base // CH) % NCTX]
            if base >= NCTX * CH:
                cx.sync()                            # the chunk this context handled one round ago must have drained
            for j in range(CH):
                i = (base + j) % len(caps)
                cx.upload(j, [t.numpy() for t in pin_in[i]], sync=False)
                if os.environ.get("VTMGPU_E2E_SKIP_SIDE") == "1" and ncalls[0] > 0:
                    continue                 # experiment switch (tools/abtest.sh): what the per-picture side information costs
                if dense_records:
                    cx.set_deblock(j, side[i][0], sync=False)
                else:
                    cx.set_deblock_sparse(j, side[i][0])
                cx.set_sao(j, side[i][1])
                cx.set_alf(j, side[i][2])
            cx.filter(0, CH, sync=False)
            for j in range(CH):
                cx.download(j, [t.numpy() for t in pin_out[(base // CH) % NCTX][j]], sync=False)
        issue[0] += time.perf_counter() - t_issue
        ncalls[0] += 1
        for cx in ectx:
            cx.sync()

    e2e_step()
    issue[0] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    barrier()
    e2e_val = world * B * px_per_pic / e2e_s / 1e6
    launches_total = launches + sum(c.launch_count() for c in ectx)
    # the e2e result must equal the device-resident result (spot check on one picture)
    ref_out = ctx.download(0)
    ok = all(np.array_equal(ref_out[k], pin_out[0][0][k].numpy()) for k in range(len(ref_out)))
    for c in ectx:
        c.close()

    line = {"metric": METRIC, "value": round(value, 1), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pictures_per_step_per_gpu": B, "pictures": workload,
                       "l2": "inputs of one step (%.0f MB per GPU) exceed the 126 MB L2, no flush needed" % (B * 24.9),
                       "activity": activity_summary(caps), "e2e_equals_resident": bool(ok)},
            "clocks": clocks, "roofline": roofline, "gpu_launches": int(launches),
            "e2e": {"value": round(e2e_val, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": int(h2d * B), "d2h_bytes_per_step": int(d2h * B),
                    "steps": args.e2e_steps, "gpu_launches": int(launches_total - launches),
                    "host_issue_ms_per_step": round(issue[0] * 1e3 / args.e2e_steps, 2), "numa_node_of_rank0": numa, "deblock_records": "dense arrays" if dense_records else "lists of active units", "ms_per_step": round(e2e_s * 1e3, 2)}}
    if world == 1 and not args.no_stress:
        # content-independent stress number: seeded pictures with every tool forced on in every CTU (SAO, luma / chroma
        # ALF with non-linear APS filters, CC-ALF) and dense small blocks for the deblocking, replayed device-resident
        from vvc_b200 import synth
        ns = min(16, B)
        scaps = [synth.make_picture(W4K, H4K, seed=4100 + i, density=1.0, p_split=0.9) for i in range(2)]
        for k in range(ns):
            ctx.set_capture(k, scaps[k % 2])
        ctx.sync()
        for _ in range(3):
            ctx.rewind(0, ns)
            ctx.filter(0, ns, sync=False)
        ctx.timer_start()
        for _ in range(5):
            ctx.rewind(0, ns)
            ctx.filter(0, ns, sync=False)
        ms = ctx.timer_stop() / 5
        line["stress_forced_on"] = {"value": round(ns * W4K * H4K / (ms * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "pictures_per_step": ns,
                                    "what": "seeded 4K pictures, every tool on in every CTU (vvc_b200/synth.py density 1.0, p_split 0.9), device-resident",
                                    "activity": activity_summary(scaps)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_reference_run()
        if res:
            line["cpu_baseline"] = {"value": round(res[0], 2), "unit": "Mpixel/s", "cores": res[1], "kind": res[2], "sample": res[3]}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
