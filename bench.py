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
# BASELINE config 3 is quoted on 64 frames: ra_2160p_b.bin holds frames 32..63 of the same synthetic sequence (32 pictures), ra_2160p_8.bin frames 0..7
STREAMS_4K = [os.path.join(ROOT, "tests", "golden", "streams", n) for n in ("ra_2160p_b.bin", "ra_2160p_8.bin")]
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
    streams = [f for f in STREAMS_4K if os.path.exists(f)]
    if streams and os.path.exists(DEC_GPU) and not os.environ.get("VTMGPU_BENCH_SYNTH"):
        caps, oks, total = [], 0, 0
        for stream in streams:
            if len(caps) >= distinct:
                break
            tmp = tempfile.mkdtemp(prefix="vtmgpu_cap_")
            try:
                env = dict(os.environ, VTMGPU_SHIM_BACKEND="gpu", VTMGPU_CAPTURE_DIR=tmp, VTMGPU_CAPTURE_PRE_ONLY="1", VTMGPU_DEVICE=str(device))
                if not env.get("VTMGPU_LIB"):
                    env.pop("VTMGPU_LIB", None)
                r = subprocess.run([DEC_GPU, "-b", stream, "-d", "0"], env=env, capture_output=True, text=True, timeout=900)
                ok = r.stdout.count("(OK)")
                files = sorted(glob.glob(os.path.join(tmp, "*.cap")))
                if not (r.returncode == 0 and ok == len(files) and ok >= 1 and "ERROR" not in r.stdout):
                    sys.stderr.write("bench.py: DecoderApp_gpu capture of %s failed (rc=%d, OK=%d): %s\n" % (os.path.basename(stream), r.returncode, ok, (r.stdout + r.stderr)[-400:]))
                    caps = []
                    break
                oks += ok
                total += len(files)
                caps += [capture.load(f) for f in files[:distinct - len(caps)]]
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        if caps:
            return caps, ("captured: VTM-encoded (RA, QP32, CC-ALF) synthetic 4K YUV (%s) decoded by DecoderApp_gpu, %d/%d pictures MD5 (OK), "
                          "%d distinct pictures replayed" % (" + ".join(os.path.basename(f) for f in streams), oks, total, len(caps)))
    caps = [synth.make_picture(W4K, H4K, seed=2160 + i, density=0.6) for i in range(distinct)]
    return caps, "synthetic planes + synthetic side info (vvc_b200/synth.py density 0.6), %d distinct pictures" % distinct


def activity_summary(caps):
    keys = sorted({k for c in caps for k in c.activity() if "frac" in k})
    acts = [c.activity() for c in caps]
    return {k: round(sum(a.get(k, 0.0) for a in acts) / len(acts), 3) for k in keys}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own filter classes on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_single(scalar=False):
    """ONE reference decoder process with the host otherwise idle (the single-core figure beside the all-core one), with the
    reference's SIMD ALF (AVX2) or with its plain C++ routines (= the stock decoder's --SIMD=SCALAR).  Mpixel/s of the filter stages or None."""
    dec = os.path.join(ROOT, "oracle", "_ref", "DecoderApp_cap")
    stream = os.path.join(ROOT, "tests", "golden", "streams", "ra_2160p_8.bin")
    if not (os.path.exists(dec) and os.path.exists(stream)):
        return None
    env = dict(os.environ, VTMGPU_SHIM_BACKEND="ref", VTMGPU_SHIM_TIMING="1")
    if scalar:
        env["VTMGPU_REF_SIMD"] = "SCALAR"
    r = subprocess.run([dec, "-b", stream, "-d", "0"], env=env, capture_output=True, text=True, timeout=600)
    m = re.search(r"vtmgpu-shim-timing: pictures=(\d+) luma_pixels=(\d+) filter_s=([0-9.eE+-]+)", r.stdout)
    if r.returncode != 0 or not m or "ERROR" in r.stdout:
        return None
    return round(int(m.group(2)) / float(m.group(3)) / 1e6, 2)


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
                if best is None or v > best:
                    best = v
                    cpu_reference_run.per_process = round(px / len(secs) / (sum(secs) / len(secs)) / 1e6, 2)     # what ONE decoder process gets (all cores busy)
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


def decoder_e2e_run(device):
    """The real drop-in: ONE DecoderApp_gpu process decodes the 32-picture 4K stream (the decoder's own picture buffers, page-locked by
    the shim on first use; synchronous per-picture calls; everything the host does for the deblocking records included); the shim's
    steady_clock timers bracket our three entry points.  Default shim configuration = the block structure is flattened on the host and
    the records are derived on the device (k_dbf_derive); the CU-walk variant (records derived by host threads) and the variant that
    also brings the reference-picture margins along are timed beside it.  Returns a dict or None."""
    stream = STREAMS_4K[0]
    if not (os.path.exists(stream) and os.path.exists(DEC_GPU)):
        return None
    # VTMGPU_SHIM_EXTEND=0: the timed region holds what the reference's three calls hold; the reference extends the picture border
    # elsewhere (Picture::extendPicBorder on first use as a reference)
    env = dict(os.environ, VTMGPU_SHIM_BACKEND="gpu", VTMGPU_SHIM_TIMING="1", VTMGPU_DEVICE=str(device), VTMGPU_SHIM_EXTEND="0")
    if not env.get("VTMGPU_LIB"):
        env.pop("VTMGPU_LIB", None)

    def run(**kw):
        r = subprocess.run([DEC_GPU, "-b", stream, "-d", "0"], env=dict(env, **kw), capture_output=True, text=True, timeout=900)
        m = re.search(r"vtmgpu-shim-timing: pictures=(\d+) luma_pixels=(\d+) filter_s=([0-9.eE+-]+) .* derive_s=([0-9.eE+-]+) .* device_derived=(\d+) .* flatten_s=([0-9.eE+-]+)", r.stdout)
        if r.returncode != 0 or not m or "ERROR" in r.stdout or r.stdout.count("(OK)") != int(m.group(1)):
            return None
        pics, px, filt, der, dev, flat = int(m.group(1)), int(m.group(2)), float(m.group(3)), float(m.group(4)), int(m.group(5)), float(m.group(6))
        tot = filt + der + flat
        return {"value": round(px / tot / 1e6, 1), "unit": "Mpixel/s", "pictures": pics, "seconds": round(tot, 5), "filter_calls_s": round(filt, 5),
                "host_derivation_s": round(der, 5), "host_flatten_s": round(flat, 5), "device_derived_pictures": dev, "ms_per_picture": round(tot * 1e3 / pics, 3)}

    best = None
    for _ in range(2):
        cur = run()
        if cur is None:
            return None
        if best is None or cur["seconds"] < best["seconds"]:
            best = cur
    best["what"] = ("one DecoderApp_gpu process on %s: time inside loopFilterPic + SAOProcess + ALFProcess incl. the host's share of the deblocking "
                    "derivation (flattening the block structure; the records are derived on the device), the decoder's picture buffers page-locked on "
                    "first use, synchronous per-picture calls, all pictures MD5 (OK)" % os.path.basename(stream))
    walk = run(VTMGPU_SHIM_DEVICE_DERIVE="0")
    if walk is not None:
        best["records_derived_by_host_threads"] = {k: walk[k] for k in ("seconds", "filter_calls_s", "host_derivation_s", "ms_per_picture")}
    ext = run(VTMGPU_SHIM_EXTEND="1")
    if ext is not None:
        best["with_border_extension_on_device"] = {k: ext[k] for k in ("seconds", "ms_per_picture")}
    return best


def band_leg(local, rank, world, dist, iters=200, warm=10):
    """BASELINE config 4 (world > 1): ONE VTM-encoded 7680x4320 intra picture filtered by all ranks in CTU-row bands over PEER
    MEMORY -- k_dbf_sao stores the halo rows into the neighbours' planes and releases flags, k_alf acquires them; no copies, no
    NCCL, no host synchronisation inside an iteration (vvc_b200/bands.py, vtmgpu_band_*).  Device-resident replay, CUDA events on
    the context's stream, max over ranks; every rank checks its band bit for bit against the whole picture filtered on its own GPU."""
    import glob as _glob
    import shutil
    import tempfile
    import numpy as np
    import torch
    from vvc_b200 import abi, bands, capture, gpu
    stream = os.path.join(ROOT, "tests", "golden", "streams", "ai_4320p.bin")
    if not (os.path.exists(stream) and os.path.exists(DEC_GPU)):
        return None
    tmp = tempfile.mkdtemp(prefix="vtmgpu_cap8k_")
    try:
        env = dict(os.environ, VTMGPU_SHIM_BACKEND="gpu", VTMGPU_CAPTURE_DIR=tmp, VTMGPU_CAPTURE_PRE_ONLY="1", VTMGPU_DEVICE=str(local))
        if not env.get("VTMGPU_LIB"):
            env.pop("VTMGPU_LIB", None)
        r = subprocess.run([DEC_GPU, "-b", stream, "-d", "0"], env=env, capture_output=True, text=True, timeout=900)
        files = sorted(_glob.glob(os.path.join(tmp, "*.cap")))
        if r.returncode != 0 or "(OK)" not in r.stdout or "ERROR" in r.stdout or not files:
            return None
        cap = capture.load(files[0])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    dev = torch.device("cuda", local)
    # the whole picture on this GPU: the reference for the bit check, and the single-GPU time
    one = gpu.Context(cap.seq, capacity=1, device=local)
    one.set_capture(0, cap)
    one.filter(0, 1)
    want = one.download(0)
    for _ in range(5):
        one.rewind(0, 1)
        one.filter(0, 1, sync=False)
    one.timer_start()
    for _ in range(50):
        one.rewind(0, 1)
        one.filter(0, 1, sync=False)
    ms_one = one.timer_stop() / 50
    one.close()
    ctx = gpu.Context(cap.seq, capacity=1, device=local)
    y0, y1 = bands.load_band(cap, ctx, rank, world)
    bands.connect_peers(ctx, rank, world, dist, dev)
    for _ in range(warm):
        ctx.rewind(0, 1)
        ctx.band_filter(0)
    ctx.sync()
    dist.barrier()
    ctx.timer_start()
    for _ in range(iters):
        ctx.rewind(0, 1)
        ctx.band_filter(0)
    ms = ctx.timer_stop() / iters
    t = torch.tensor([ms, ms_one], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = [np.zeros_like(p) for p in cap.pre]
    ctx.download_rows(0, out, y0, y1)
    sy = abi.chroma_shifts(cap.seq["chroma_format"])[1]
    ok = all(np.array_equal(out[c][y0 >> (sy if c else 0):y1 >> (sy if c else 0)], want[c][y0 >> (sy if c else 0):y1 >> (sy if c else 0)]) for c in range(cap.ncomp))
    oks = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    dist.barrier()
    ctx.band_disconnect()
    ctx.close()
    ms, ms_one = float(t[0].item()), float(t[1].item())
    px = cap.width * cap.height
    return {"metric": "DBF+SAO+ALF Mpixel/s, one 7680x4320 intra picture in CTU-row bands", "value": round(px / (ms * 1e-3) / 1e6, 1), "unit": "Mpixel/s",
            "n_gpus": world, "iterations": iters, "ms_per_picture": round(ms, 4), "single_gpu_ms_per_picture": round(ms_one, 4), "speedup_vs_single_gpu": round(ms_one / ms, 2),
            "scaling": "strong", "bands": bands.band_rows(cap.height, world), "bit_exact_vs_single_gpu": bool(oks.item()),
            "halo": "4 rows per plane and border stored into the neighbour's plane by k_dbf_sao (NVLink peer stores) + release / acquire flags; no copies, no NCCL, no host sync per iteration"}


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
    ap.add_argument("--distinct", type=int, default=40, help="distinct pictures (the two committed 4K streams hold 32 + 8) replicated to fill the batch")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the forced-on synthetic leg (N=1 only)")
    ap.add_argument("--no-decoder", action="store_true", help="skip the in-decoder leg (N=1 only)")
    ap.add_argument("--no-bands", action="store_true", help="skip the 8K band leg (N>1 only)")
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
            ent = json.load(open(tr)).get(roofline["kernel"], {})
            per_px = ent.get("dram_bytes_per_luma_pixel")
            if per_px is not None:
                roofline["traffic"] = round(per_px * px_per_pic * B)      # ncu dram__bytes_read+write per luma pixel x pixels of one launch
                same = ent.get("pictures_in_launch") == B
                roofline["traffic_source"] = "%s: ncu --set full of one %s-picture launch of this loop (profiles/r2_ncu_full_summary.csv, tools/profile_round.sh), dram__bytes_read + write" % (
                    "measured per launch" if same else "extrapolated (bytes per luma pixel x the pixels of this launch)", ent.get("pictures_in_launch", "?"))
                roofline["traffic_over_algorithmic"] = round(per_px / (B_ALG_SAOALF if dom else B_ALG_DBF), 3)
        except Exception:
            pass

    # ---- end to end through the C ABI with host buffers -------------------------------------------------------------
    # ONE C call per step (vtmgpu_batch_filter): the library walks the pictures round robin over LANES single-picture contexts
    # (the pictures in flight) -- pinned H2D of the planes, record lists, SAO / ALF parameters on an upload stream, the chain on a
    # second stream, pinned D2H on a third, ordered by events -- issued by this thread
    import ctypes as C
    from vvc_b200 import abi
    LANES = int(os.environ.get("VTMGPU_E2E_LANES", "8"))
    batch = gpu.Batch(seq, lanes=LANES, device=local)
    pin_in = [[torch.from_numpy(p.copy()).pin_memory() for p in c.pre] for c in caps]
    pin_out = [[torch.empty_like(t).pin_memory() for t in pin_in[0]] for _ in range(LANES)]
    side = []
    for c in caps:
        ctus = c.sao_ctus()
        if ctus is not None:
            gpu.sao_reconstruct(ctus, c.width_in_ctus, c.ncomp, c.sao_scale[0], c.sao_scale[1])
        # the decoder-side producer of the segment records (the shim's CU walk) appends the active units to lists in
        # page-locked memory: no staging copy in the library, ~0.15 instead of 0.75 B of records per luma pixel on the bus
        dp = gpu.sparse_records(c.dbf_luma, c.dbf_chroma if c.ncomp > 1 else None, pin=True)
        nbytes = sum(dp.luma_count[d] * C.sizeof(abi.DbfLumaEntry) + dp.chroma_count[d] * C.sizeof(abi.DbfChromaEntry) for d in range(2))
        side.append((dp, ctus, c.alf_params(), nbytes, c.vb_struct()))
    plane_bytes = sum(t.numel() * 2 for t in pin_in[0])
    h2d = sum(plane_bytes + side[k % len(caps)][3] for k in range(B)) / B       # per picture, averaged over the batch
    d2h = plane_bytes
    pics = [gpu.host_picture([t.numpy() for t in pin_in[k % len(caps)]], [t.numpy() for t in pin_out[k % LANES]], side[k % len(caps)][0],
                             side[k % len(caps)][1], side[k % len(caps)][2], side[k % len(caps)][4]) for k in range(B)]
    arr = (abi.HostPicture * B)(*pics)

    batch.filter(arr)                                 # warm-up (allocations of the record landing areas, first-touch)
    e2e_launch0 = batch.launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        batch.filter(arr)                             # returns when every output has landed in host memory
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    barrier()
    e2e_val = world * B * px_per_pic / e2e_s / 1e6
    e2e_launches = batch.launch_count() - e2e_launch0
    # the e2e result must equal the device-resident result: the last picture of lane 0 against the same picture's slot
    last = max(k for k in range(B) if k % LANES == 0)
    pic = last % len(caps)
    ref_out = ctx.download(pic)                  # resident slot s holds picture s % len(caps), and pic < min(B, len(caps))
    ok = all(np.array_equal(ref_out[k], pin_out[0][k].numpy()) for k in range(len(ref_out)))
    batch.close()
    if not ok:
        raise SystemExit("bench.py: the end-to-end output of picture %d differs from the device-resident output -- no number is reported" % pic)

    # what the host <-> device path of THIS box allows for these byte counts: the same bytes per picture as plain concurrent copies
    # (one H2D stream, one D2H stream, page-locked, nothing else running) -- all ranks at once, like the e2e leg
    dev_in = torch.empty(int(h2d) // 2, dtype=torch.int16, device="cuda")
    dev_out = torch.empty(int(d2h) // 2, dtype=torch.int16, device="cuda")
    host_in = torch.empty(int(h2d) // 2, dtype=torch.int16).pin_memory()
    host_out = torch.empty(int(d2h) // 2, dtype=torch.int16).pin_memory()
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def copy_step():
        for _ in range(B):
            with torch.cuda.stream(s_up):
                dev_in.copy_(host_in, non_blocking=True)
            with torch.cuda.stream(s_dn):
                host_out.copy_(dev_out, non_blocking=True)
        s_up.synchronize()
        s_dn.synchronize()

    copy_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        copy_step()
    copy_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    barrier()
    ceiling = world * B * px_per_pic / copy_s / 1e6
    del dev_in, dev_out, host_in, host_out

    line = {"metric": METRIC, "value": round(value, 1), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pictures_per_step_per_gpu": B, "pictures": workload,
                       "l2": "inputs of one step (%.0f MB per GPU) exceed the 126 MB L2, no flush needed" % (B * 24.9),
                       "activity": activity_summary(caps), "e2e_equals_resident": bool(ok)},
            "clocks": clocks, "roofline": roofline, "gpu_launches": int(launches),
            "e2e": {"value": round(e2e_val, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": int(h2d * B), "d2h_bytes_per_step": int(d2h * B),
                    "steps": args.e2e_steps, "gpu_launches": int(e2e_launches // max(1, args.e2e_steps)),
                    "api": "vtmgpu_batch_filter: one C call per step, %d lanes (pictures in flight), three streams (uploads / kernels / downloads), one issuing thread per GPU" % LANES,
                    "numa_node_of_rank0": numa, "deblock_records": "lists of active units", "ms_per_step": round(e2e_s * 1e3, 2),
                    "pcie_ceiling": {"value": round(ceiling, 1), "unit": "Mpixel/s", "ms_per_step": round(copy_s * 1e3, 2),
                                     "h2d_gbs": round(h2d * B / copy_s / 1e9, 1), "d2h_gbs": round(d2h * B / copy_s / 1e9, 1),
                                     "what": "the same bytes per picture as plain concurrent pinned copies (one H2D + one D2H stream per rank, all ranks at once), measured in this run"},
                    "frac_of_pcie_ceiling": round(e2e_val / ceiling, 3)}}
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
                                    "chain_frac": round(B_ALG_CHAIN * ns * W4K * H4K / (ms * 1e-3) / 1e9 / peak, 4),
                                    "what": "seeded 4K pictures, every tool on in every CTU (vvc_b200/synth.py density 1.0, p_split 0.9), device-resident",
                                    "activity": activity_summary(scaps)}
    if world > 1 and not args.no_bands:
        try:
            res = band_leg(local, rank, world, dist)
        except Exception as e:                       # the picture-parallel line must survive a box without peer access
            res = {"error": repr(e)[:300]}
        if res is not None:
            line["band_8k"] = res
    if rank == 0 and world == 1 and not args.no_decoder:
        # the drop-in as a decoder user meets it (host derivation, pageable buffers, synchronous calls); bench "e2e" above replays
        # pre-derived side information from pinned memory through pipelined contexts -- both are reported, neither hides the other
        dec = decoder_e2e_run(local)
        if dec:
            line["e2e_decoder"] = dec
            line["e2e"]["host_derivation_excluded"] = True
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_reference_run()
        if res:
            line["cpu_baseline"] = {"value": round(res[0], 2), "unit": "Mpixel/s", "cores": res[1], "kind": res[2], "sample": res[3]}
            if getattr(cpu_reference_run, "per_process", None):
                line["cpu_baseline"]["per_process"] = cpu_reference_run.per_process
            if res[2] == "reference":
                # one process alone on the host (SIMD and plain C++ ALF) beside the all-core aggregate
                one, scalar = cpu_reference_single(False), cpu_reference_single(True)
                if one is not None:
                    line["cpu_baseline"]["single_process_alone"] = one
                if scalar is not None:
                    line["cpu_baseline"]["single_process_alone_scalar_alf"] = scalar
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
