"""Band mode (one picture over several ranks).

CPU: the partition, the halo plan and the exchange choreography run with world_size 2 over gloo; the per-band filter is
the oracle working on a picture whose rows outside band +- 4 are poisoned, so the test also proves that 4 rows of SAO
output are all the ALF stage needs across a band border.
GPU: the CUDA path in bands (several contexts, set_rows / export / import) == the oracle on the whole picture."""
import os
import socket
import sys

import numpy as np
import pytest

import pyoracle
from vvc_b200 import bands, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_rows_partition():
    assert bands.band_rows(4320, 8) == [(0, 640), (640, 1280), (1280, 1792), (1792, 2304), (2304, 2816), (2816, 3328), (3328, 3840), (3840, 4320)]
    assert bands.band_rows(2160, 2) == [(0, 1152), (1152, 2160)]
    assert bands.band_rows(240, 2) == [(0, 128), (128, 240)]
    assert bands.band_rows(128, 1) == [(0, 128)]
    with pytest.raises(ValueError):
        bands.band_rows(240, 3)
    for h, n in ((4320, 8), (1080, 4), (2160, 3)):
        b = bands.band_rows(h, n)
        assert b[0][0] == 0 and b[-1][1] == h and all(b[i][1] == b[i + 1][0] for i in range(n - 1))
        assert all(y0 % 128 == 0 for y0, _ in b)


def test_halo_plan_is_symmetric():
    b = bands.band_rows(1080, 4)
    shifts = [0, 1, 1]
    for r in range(4):
        for peer, kind, comp, row, n in bands.halo_plan(b, r, shifts):
            other = "recv" if kind == "send" else "send"
            mirror = [m for m in bands.halo_plan(b, peer, shifts) if m[0] == r and m[1] == other and m[2] == comp]
            assert len(mirror) == 1 and mirror[0][4] == n
            if kind == "send":                      # what one rank sends is exactly what the other receives
                assert mirror[0][3] == row


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_rank(rank, world, port, seed, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        cap = synth.make_picture(192, 384, seed=seed, density=0.9)
        ref = pyoracle.filter_capture(cap)                       # whole picture: after SAO, after ALF
        b = bands.band_rows(cap.height, world)
        y0, y1 = b[rank]
        # this rank's SAO output: only its band is valid, everything else poisoned
        mine = [np.full_like(p, 777) for p in ref["sao"]]
        for c, p in enumerate(ref["sao"]):
            sy = 0 if c == 0 else 1
            mine[c][y0 >> sy:y1 >> sy] = p[y0 >> sy:y1 >> sy]
        widths = [p.shape[1] for p in mine]

        def export_fn(comp, row, n, buf):
            buf.copy_(torch.from_numpy(mine[comp][row:row + n].reshape(-1).copy()))

        def import_fn(comp, row, n, buf):
            mine[comp][row:row + n] = buf.numpy().reshape(n, widths[comp])

        bands.exchange(bands.halo_plan(b, rank, [0, 1, 1]), export_fn, import_fn, dist, lambda c: widths[c],
                       lambda n: torch.empty(n, dtype=torch.int16))
        pyoracle.alf(cap.seq, mine, cap.alf_params())            # in place
        got = mine
        ok = all(np.array_equal(got[c][y0 >> (0 if c == 0 else 1):y1 >> (0 if c == 0 else 1)],
                                ref["final"][c][y0 >> (0 if c == 0 else 1):y1 >> (0 if c == 0 else 1)]) for c in range(3))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_band_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_rank, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)], res


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,feat", [(416, 384, 2, False), (256, 640, 3, False), (1920, 1080, 4, False), (512, 640, 3, True)])
def test_bands_on_one_gpu_match_oracle(w, h, n, feat):
    # feat: slice / tile clip flags and signalled virtual boundaries (one of them 8 rows above a band border) in band mode
    cap = synth.make_picture(w, h, seed=w + n, density=0.8, partitions=feat, vb=([200], [248, 504]) if feat else None)
    want = pyoracle.filter_capture(cap)["final"]
    got = bands.filter_picture_in_bands_local(cap, n)
    for c in range(3):
        bad = np.argwhere(got[c] != want[c])
        assert len(bad) == 0, "component %d differs at %d samples, first %s" % (c, len(bad), tuple(bad[0]))


def _gpu_rank(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from vvc_b200 import gpu
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cap = synth.make_picture(1920, 1080, seed=5, density=0.8)
        want = pyoracle.filter_capture(cap)["final"]
        ctx = gpu.Context(cap.seq, capacity=1, device=rank)
        out, (y0, y1) = bands.filter_picture_in_bands(cap, ctx, rank, world, dist)
        ctx.close()
        ok = all(np.array_equal(out[c][y0 >> (0 if c == 0 else 1):y1 >> (0 if c == 0 else 1)],
                                want[c][y0 >> (0 if c == 0 else 1):y1 >> (0 if c == 0 else 1)]) for c in range(3))
        q.put((rank, bool(ok)))
    except Exception as e:                                         # report instead of letting the parent time out
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _gpu_rank_peer(rank, world, port, q):
    """Peer band mode: several iterations back to back (the flag protocol must order iteration i+1's stores after the
    neighbour's reads of iteration i), a picture with every stage on and one with SAO / ALF off (the kernels must still signal)."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from vvc_b200 import gpu
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ok = True
        for seed, density, iters in ((5, 0.8, 7), (9, 0.0, 3), (11, 1.0, 2)):
            cap = synth.make_picture(1920, 1080, seed=seed, density=density)
            want = pyoracle.filter_capture(cap)["final"]
            ctx = gpu.Context(cap.seq, capacity=1, device=rank)
            out, (y0, y1) = bands.filter_picture_in_bands_peer(cap, ctx, rank, world, dist, iterations=iters)
            ctx.close()
            ok = ok and all(np.array_equal(out[c][y0 >> (0 if c == 0 else 1):y1 >> (0 if c == 0 else 1)],
                                           want[c][y0 >> (0 if c == 0 else 1):y1 >> (0 if c == 0 else 1)]) for c in range(3))
        q.put((rank, bool(ok)))
    except Exception as e:                                         # report instead of letting the parent time out
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_bands_peer_memory(world):
    """The band pipeline over peer memory (halo rows stored into the neighbours' planes by k_dbf_sao, flags acquired by k_alf):
    bit-exact against the oracle on `world` GPUs, one process per GPU."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_rank_peer, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(r, True) for r in range(world)], res


@pytest.mark.gpu
def test_bands_two_gpus_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)], res
