"""Multi-GPU frame sharding: one process per GPU, CPI frames are independent units.

The reference parallelises over CPI frames implicitly (one dask chunk = one frame on a
thread pool, ``main.py:169-194``); nothing couples frames on the hot path.  So the N-GPU
layout is: frame ``f`` of the stream belongs to rank ``f % world`` (or a contiguous range),
each rank runs its own ``FramePipeline``, and the only data that ever moves between ranks is
*input staging* when a single ingest rank holds the IQ stream:

* ``scatter_frames``      ingest rank sends each rank its own (ref, srv) frames (1/G of the data
                          per link) -- ``torch.distributed.scatter``, NCCL over NVLink on GPUs
* ``broadcast_reference`` all ranks need the same reference-channel block (one illuminator,
                          several surveillance channels / overlapped windows) -- one NCCL broadcast
* ``gather_maps``         optional: range-Doppler maps (0.6 MB/frame) back to the ingest rank

There is no reduction, no all-to-all and no collective inside the timed compute path; when
every rank reads or synthesises its own frames there is no communication at all.  The same
code runs on CPU tensors with the ``gloo`` backend (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import numpy as np


def shard_indices(nframes: int, rank: int, world: int, mode: str = "interleaved") -> np.ndarray:
    """Frame indices owned by ``rank``.  ``interleaved``: f % world == rank (keeps ranks in
    lock-step on a live stream); ``contiguous``: equal consecutive ranges (overlapped CPIs
    re-read only a half-CPI halo)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if mode == "interleaved":
        return np.arange(rank, nframes, world)
    if mode == "contiguous":
        base, extra = divmod(nframes, world)
        start = rank * base + min(rank, extra)
        return np.arange(start, start + base + (1 if rank < extra else 0))
    raise ValueError(f"unknown sharding mode {mode!r}")


def _as_real(t):
    import torch
    return torch.view_as_real(t) if t.is_complex() else t


def scatter_frames(ref_frames, srv_frames, nframes, n, src=0, device=None, mode="interleaved", group=None):
    """Ingest rank ``src`` holds ``(nframes, n)`` complex64 tensors; every rank returns its shard
    ``(ref_local, srv_local, indices)``.  Other ranks pass ``None`` for the frames."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    idx = shard_indices(nframes, rank, world, mode)
    counts = [len(shard_indices(nframes, r, world, mode)) for r in range(world)]
    width = max(counts) if counts else 0
    dev = device if device is not None else (ref_frames.device if ref_frames is not None else torch.device("cpu"))
    out = []
    for frames in (ref_frames, srv_frames):
        recv = torch.zeros((width, n), dtype=torch.complex64, device=dev)
        chunks = None
        if rank == src:
            chunks = []
            for r in range(world):
                ids = torch.as_tensor(shard_indices(nframes, r, world, mode), device=frames.device)
                c = torch.zeros((width, n), dtype=torch.complex64, device=dev)
                if len(ids):
                    c[: len(ids)] = frames.index_select(0, ids).to(dev)
                chunks.append(_as_real(c).contiguous())
        dist.scatter(_as_real(recv), chunks, src=src, group=group)
        out.append(recv[: len(idx)])
    return out[0], out[1], idx


def broadcast_reference(ref_block, src=0, group=None):
    """In-place broadcast of the reference-channel block (complex64 tensor) from ``src``."""
    import torch.distributed as dist
    dist.broadcast(_as_real(ref_block), src=src, group=group)
    return ref_block


def gather_maps(maps_local, nframes, src=0, mode="interleaved", group=None):
    """Collect per-rank maps ``(nlocal, F, R+1)`` on ``src`` in stream order; others get ``None``."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = [len(shard_indices(nframes, r, world, mode)) for r in range(world)]
    width = max(counts) if counts else 0
    shape = (width,) + tuple(maps_local.shape[1:])
    send = torch.zeros(shape, dtype=maps_local.dtype, device=maps_local.device)
    send[: maps_local.shape[0]] = maps_local
    bufs = None
    if rank == src:
        bufs = [torch.zeros_like(_as_real(send)) for _ in range(world)]
    dist.gather(_as_real(send).contiguous(), bufs, dst=src, group=group)
    if rank != src:
        return None
    full = torch.zeros((nframes,) + tuple(maps_local.shape[1:]), dtype=maps_local.dtype, device=maps_local.device)
    for r in range(world):
        ids = shard_indices(nframes, r, world, mode)
        if len(ids):
            got = torch.view_as_complex(bufs[r]) if maps_local.is_complex() else bufs[r]
            full[torch.as_tensor(ids, device=full.device)] = got[: len(ids)]
    return full


def process_stream(process_local, ref_frames, srv_frames, nframes, n, src=0, device=None,
                   mode="interleaved", gather=True, group=None):
    """scatter -> ``process_local(ref_local, srv_local) -> maps_local`` -> (optional) gather.

    ``process_local`` is ``lambda r, s: pipe.run_device(r, s, out)`` on a GPU rank; the CPU
    tests pass a stand-in so the plumbing is exercised with gloo."""
    ref_l, srv_l, idx = scatter_frames(ref_frames, srv_frames, nframes, n, src, device, mode, group)
    maps_l = process_local(ref_l, srv_l)
    if not gather:
        return maps_l, idx
    return gather_maps(maps_l, nframes, src, mode, group), idx


def stream_benchmark(pipe, ref_d, srv_d, maps_d, nframes_total, chunk, rank, world, device, src=0):
    """BASELINE config 3 with real staging: a stream of ``nframes_total`` CPI frames is held by the ingest rank ``src``
    (its resident frames stand in for the stream), frame ``f`` belongs to rank ``f % world``; the ingest rank sends every
    other rank its (ref, srv) frames in chunks of ``chunk`` frames over NCCL point-to-point (NVLink / NVSwitch) on a side
    stream, double buffered against the rank's compute of the previous chunk; results stay on the rank.  This is the
    traffic the dask graph's halo staging (main.py:178-181) turns into when the chunks live on different GPUs.
    Returns (rank 0) a dict: whole-job frames/s including the staging, bytes the ingest rank sent, its egress GB/s.
    Time = max over ranks of the device time between the first enqueue and the last result."""
    import torch
    import torch.distributed as dist
    n = ref_d.shape[1]
    res = ref_d.shape[0]
    per_rank = nframes_total // world
    nchunks = -(-per_rank // chunk)
    side = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    bufs = None
    if rank != src:
        bufs = [(torch.empty((chunk, n), dtype=torch.complex64, device=device), torch.empty((chunk, n), dtype=torch.complex64, device=device))
                for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]       # staging of buffer k complete
    freed = [torch.cuda.Event() for _ in range(2)]       # compute on buffer k complete

    def run_once():
        for c in range(nchunks + 1):
            m = min(chunk, per_rank - c * chunk) if c < nchunks else 0
            # ---- staging of chunk c (side stream)
            if c < nchunks:
                with torch.cuda.stream(side):
                    if rank == src:
                        ops = []
                        for r in range(world):
                            if r == src:
                                continue
                            o = ((c * world + r) * chunk) % max(1, res - chunk + 1)
                            ops.append(dist.P2POp(dist.isend, _as_real(ref_d[o:o + m]), r))
                            ops.append(dist.P2POp(dist.isend, _as_real(srv_d[o:o + m]), r))
                        for w in dist.batch_isend_irecv(ops):
                            w.wait()
                    else:
                        k = c & 1
                        if c >= 2:
                            side.wait_event(freed[k])
                        rb, sb = bufs[k]
                        ops = [dist.P2POp(dist.irecv, _as_real(rb[:m]), src), dist.P2POp(dist.irecv, _as_real(sb[:m]), src)]
                        for w in dist.batch_isend_irecv(ops):
                            w.wait()
                        ready[k].record(side)
            # ---- compute of chunk c - 1 (main stream)
            if c >= 1:
                cc = c - 1
                mm = min(chunk, per_rank - cc * chunk)
                o = (cc * chunk) % max(1, res - chunk + 1)
                if rank == src:
                    pipe.run_device(ref_d[o:o + mm], srv_d[o:o + mm], maps_d[o:o + mm])
                else:
                    k = cc & 1
                    main.wait_event(ready[k])
                    rb, sb = bufs[k]
                    pipe.run_device(rb[:mm], sb[:mm], maps_d[o:o + mm])
                    freed[k].record(main)
        main.wait_stream(side)

    run_once()                                           # warm-up (NCCL channels, workspaces)
    torch.cuda.synchronize(device)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    run_once()
    e1.record(main)
    torch.cuda.synchronize(device)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank != 0:
        return None
    sent = (world - 1) * per_rank * 2 * n * 8
    return {"frames": per_rank * world, "value": per_rank * world / (ms * 1e-3), "unit": "frames/s", "ms": round(ms, 3),
            "chunk_frames": chunk, "ingest_rank_sent_bytes": sent, "ingest_egress_GBps": round(sent / (ms * 1e-3) / 1e9, 1),
            "staging": "NCCL point-to-point (batch_isend_irecv) from rank 0 on a side stream, double buffered against compute",
            "note": "frame f -> rank f % world; rank 0 holds the stream in HBM and also computes its own share; the limiter is "
                    "rank 0's NVLink egress when ingest_egress_GBps approaches the ~770 GB/s peer-copy figure"}
