"""Band mode: ONE large picture (BASELINE config 4: 7680x4320) filtered by several GPUs, one CTU-row band each
(SURVEY.md 8e).

Every rank creates a context for the FULL picture geometry (so CTU positions, ALF virtual boundaries and picture-border
rules keep their absolute coordinates -- HBM is not the constraint: an 8K 4:2:0 slot is 3 x 100 MB), uploads only its
band plus the rows the deblocking of the border tiles reads, and runs

    deblocking + SAO  (exact for the band: both read pre-filter samples only)
    halo exchange     4 rows of SAO output on each side of every band border go to the neighbour rank
    ALF / CC-ALF      (reads 3 rows across the border: filter taps and the Laplacian classifier)

The exchange is the only data-path communication: a neighbour send/recv of width x 4 samples per plane (60 KB of luma
at 8K), carried by torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU test of the choreography).  The plan
(`band_rows`, `halo_plan`) is pure host logic and is shared by the GPU path and the CPU test.
"""
import numpy as np

from . import abi

HALO = 4            # rows of SAO output the ALF stage reads across a band border (box of the tile loads: +-4)
UPLOAD_MARGIN = 16  # pre-filter luma rows the deblocking of a band's border tiles reads (8 chroma rows at 4:2:0)
BAND_ALIGN = 128    # band borders sit on multiples of 128 luma rows (vtmgpu_set_rows)


def band_rows(height, world, align=BAND_ALIGN):
    """Luma row ranges [(y0, y1)] of the `world` bands: whole units of `align` rows, the larger bands first
    (4320 rows over 8 ranks -> CTU rows 5,5,4,4,4,4,4,4)."""
    units = (height + align - 1) // align
    if world > units:
        raise ValueError("more ranks (%d) than %d-row units (%d)" % (world, align, units))
    base, extra = divmod(units, world)
    out, u = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((u * align, min((u + n) * align, height)))
        u += n
    return out


def halo_plan(bands, rank, shifts):
    """Messages of one rank: list of (peer, 'send'|'recv', comp, first_row, nrows) in plane rows of component comp.
    shifts[comp] = vertical subsampling shift of the component.  Sends are the band's own outermost HALO rows, receives
    land just outside the band."""
    y0, y1 = bands[rank]
    plan = []
    for comp, sy in enumerate(shifts):
        if rank > 0:                               # upper neighbour
            plan.append((rank - 1, "send", comp, y0 >> sy, HALO))
            plan.append((rank - 1, "recv", comp, (y0 >> sy) - HALO, HALO))
        if rank + 1 < len(bands):                  # lower neighbour
            plan.append((rank + 1, "send", comp, (y1 >> sy) - HALO, HALO))
            plan.append((rank + 1, "recv", comp, y1 >> sy, HALO))
    return plan


def halo_plan_packed(bands, rank, shifts):
    """The same exchange with ONE message per neighbour and direction: list of (peer, 'send'|'recv', [first row per component])."""
    y0, y1 = bands[rank]
    plan = []
    if rank > 0:
        plan.append((rank - 1, "send", [y0 >> s for s in shifts]))
        plan.append((rank - 1, "recv", [(y0 >> s) - HALO for s in shifts]))
    if rank + 1 < len(bands):
        plan.append((rank + 1, "send", [(y1 >> s) - HALO for s in shifts]))
        plan.append((rank + 1, "recv", [y1 >> s for s in shifts]))
    return plan


def exchange_packed(plan, ctx, dist, buffers, host_sync=True):
    """Packed halo exchange on a gpu.Context: buffers[i] = device int16 tensor of (sum of plane widths) * HALO samples for
    plan entry i (allocated once by the caller and reused)."""
    ops, recvs = [], []
    for (peer, kind, rows), buf in zip(plan, buffers):
        if kind == "send":
            ctx.export_halo(0, rows, HALO, buf.data_ptr())
            ops.append(dist.P2POp(dist.isend, _wire(buf), peer))
        else:
            ops.append(dist.P2POp(dist.irecv, _wire(buf), peer))
            recvs.append((rows, buf))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if host_sync:
            import torch
            torch.cuda.synchronize()
    for rows, buf in recvs:
        ctx.import_halo(0, rows, HALO, buf.data_ptr())


def _wire(buf):
    """NCCL has no int16: samples travel as bytes (same memory)."""
    import torch
    return buf.view(torch.uint8)


def exchange(plan, export_fn, import_fn, dist, width_of, make_buffer, host_sync=True):
    """Runs a halo plan: export_fn(comp, row, n, buf), import_fn(comp, row, n, buf); buffers from make_buffer(n_samples).
    Non-blocking sends/receives towards both neighbours, then the imports.  host_sync=False when the context runs on
    torch's current stream (Context.set_stream): NCCL then orders with the exports / imports on the device."""
    ops, recvs, keep = [], [], []
    for peer, kind, comp, row, n in plan:
        buf = make_buffer(width_of(comp) * n)
        keep.append(buf)
        if kind == "send":
            export_fn(comp, row, n, buf)
            ops.append(dist.P2POp(dist.isend, _wire(buf), peer))
        else:
            ops.append(dist.P2POp(dist.irecv, _wire(buf), peer))
            recvs.append((comp, row, n, buf))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if keep[0].is_cuda and host_sync:
            # NCCL work completes on torch's stream; the imports run on the context's own stream
            import torch
            torch.cuda.synchronize()
    for comp, row, n, buf in recvs:
        import_fn(comp, row, n, buf)


def connect_peers(ctx, rank, world, dist, device):
    """Peer band mode: every rank exports the handle of its plane memory / flag block, the handles travel by all_gather, and every
    rank maps the memory of the ranks above and below (CUDA IPC -> NVLink peer access).  Call after ctx.set_rows()."""
    import torch
    mine = torch.tensor(list(ctx.band_export()), dtype=torch.uint8, device=device)
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    hs = [bytes(t.cpu().numpy().tobytes()) for t in every]
    ctx.band_connect(hs[rank - 1] if rank > 0 else None, hs[rank + 1] if rank + 1 < world else None)
    dist.barrier()


def load_band(cap, ctx, rank, world):
    """Uploads this rank's band of the captured picture (plus the rows the deblocking of its border tiles reads) and all side
    information into slot 0 of a context created for the full geometry.  Returns (y0, y1)."""
    from . import gpu
    h = cap.height
    y0, y1 = band_rows(h, world)[rank]
    ctx.set_rows(y0, y1)
    ctx.upload_rows(0, cap.pre, max(0, y0 - UPLOAD_MARGIN), min(h, y1 + UPLOAD_MARGIN))
    ctx.set_deblock(0, cap.deblock_params())
    ctus = cap.sao_ctus()
    if ctus is not None:
        gpu.sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, cap.sao_scale[0], cap.sao_scale[1])
    ctx.set_sao(0, ctus, cap.vb_struct())
    ctx.set_alf(0, cap.alf_params())
    return y0, y1


def filter_picture_in_bands_peer(cap, ctx, rank, world, dist, out=None, iterations=1):
    """The band pipeline over peer memory: one call per iteration (vtmgpu_band_filter_async), the halo rows are stored into the
    neighbours' planes by the deblocking + SAO kernel itself and the ALF kernel waits for the neighbours' flags -- no copies, no
    NCCL, no host synchronisation inside an iteration.  Returns (planes, (y0, y1)) like filter_picture_in_bands."""
    import torch
    y0, y1 = load_band(cap, ctx, rank, world)
    connect_peers(ctx, rank, world, dist, torch.device("cuda", ctx.device))
    for _ in range(iterations):
        ctx.rewind(0, 1)
        ctx.band_filter(0)
    ctx.sync()
    if out is None:
        out = [np.zeros_like(p) for p in cap.pre]
    ctx.download_rows(0, out, y0, y1)
    dist.barrier()                       # nobody unmaps while a neighbour may still store into its planes
    ctx.band_disconnect()
    return out, (y0, y1)


def filter_picture_in_bands(cap, ctx, rank, world, dist, out=None):
    """DBF -> SAO -> [halo exchange] -> ALF of this rank's band of the captured picture `cap` on context `ctx` (slot 0,
    created for the full geometry).  Returns (planes, (y0, y1)): full-size host planes whose rows [y0, y1) hold the result."""
    import torch
    from . import gpu
    h = cap.height
    bands = band_rows(h, world)
    y0, y1 = bands[rank]
    sx, sy = abi.chroma_shifts(cap.seq["chroma_format"])
    shifts = [0] + ([sy, sy] if cap.ncomp == 3 else [])
    ctx.set_rows(y0, y1)
    ctx.upload_rows(0, cap.pre, max(0, y0 - UPLOAD_MARGIN), min(h, y1 + UPLOAD_MARGIN))
    ctx.set_deblock(0, cap.deblock_params())
    ctus = cap.sao_ctus()
    if ctus is not None:
        gpu.sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, cap.sao_scale[0], cap.sao_scale[1])
    ctx.set_sao(0, ctus, cap.vb_struct())
    ctx.set_alf(0, cap.alf_params())
    ctx.deblock_sao(0, 1)                      # one kernel: the SAO of the band's border rows sees the deblocked rows across the border
    dev = torch.device("cuda", ctx.device)
    widths = [cap.width] + [cap.width >> sx] * (cap.ncomp - 1)
    plan = halo_plan_packed(bands, rank, shifts)
    exchange_packed(plan, ctx, dist, [torch.empty(sum(widths) * HALO, dtype=torch.int16, device=dev) for _ in plan])
    ctx.alf(0, 1)
    if out is None:
        out = [np.zeros_like(p) for p in cap.pre]
    ctx.download_rows(0, out, y0, y1)
    return out, (y0, y1)


def filter_picture_in_bands_local(cap, n_bands, device=0):
    """The same band pipeline driven by ONE process on ONE GPU: n_bands contexts (all for the full geometry), the halo
    rows travel through device buffers instead of NCCL.  Exercises set_rows / upload_rows / export / import / download_rows
    exactly as the multi-GPU path does; returns the assembled full picture."""
    import torch
    from . import gpu
    h = cap.height
    bands = band_rows(h, n_bands)
    sx, sy = abi.chroma_shifts(cap.seq["chroma_format"])
    shifts = [0] + ([sy, sy] if cap.ncomp == 3 else [])
    widths = [cap.width] + [cap.width >> sx] * (cap.ncomp - 1)
    dev = torch.device("cuda", device)
    ctxs = [gpu.Context(cap.seq, capacity=1, device=device) for _ in range(n_bands)]
    out = [np.zeros_like(p) for p in cap.pre]
    try:
        for r, ctx in enumerate(ctxs):
            y0, y1 = bands[r]
            ctx.set_rows(y0, y1)
            ctx.upload_rows(0, cap.pre, max(0, y0 - UPLOAD_MARGIN), min(h, y1 + UPLOAD_MARGIN))
            ctx.set_deblock(0, cap.deblock_params())
            ctus = cap.sao_ctus()
            if ctus is not None:
                gpu.sao_reconstruct(ctus, cap.width_in_ctus, cap.ncomp, cap.sao_scale[0], cap.sao_scale[1])
            ctx.set_sao(0, ctus, cap.vb_struct())
            ctx.set_alf(0, cap.alf_params())
            ctx.deblock_sao(0, 1)
        mail = {}
        for r, ctx in enumerate(ctxs):
            for peer, kind, comp, row, n in halo_plan(bands, r, shifts):
                if kind == "send":
                    buf = torch.empty(widths[comp] * n, dtype=torch.int16, device=dev)
                    ctx.export_rows(0, comp, row, n, buf.data_ptr())
                    mail[(r, peer, comp)] = buf
        for r, ctx in enumerate(ctxs):
            for peer, kind, comp, row, n in halo_plan(bands, r, shifts):
                if kind == "recv":
                    ctx.import_rows(0, comp, row, n, mail[(peer, r, comp)].data_ptr())
        for r, ctx in enumerate(ctxs):
            ctx.alf(0, 1)
            ctx.download_rows(0, out, bands[r][0], bands[r][1])
    finally:
        for ctx in ctxs:
            ctx.close()
    return out
