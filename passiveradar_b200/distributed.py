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
