"""world_size-2 gloo tests (CPU) of the frame-sharding plumbing used by the N>1 GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from passiveradar_b200 import distributed as D


def test_shard_indices_partition_the_stream():
    for nframes in (0, 1, 7, 16, 1000):
        for world in (1, 2, 3, 8):
            for mode in ("interleaved", "contiguous"):
                parts = [D.shard_indices(nframes, r, world, mode) for r in range(world)]
                allidx = np.sort(np.concatenate(parts)) if parts else np.array([])
                assert np.array_equal(allidx, np.arange(nframes))
                sizes = [len(p) for p in parts]
                assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_indices(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nframes, n, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ref = srv = None
        if rank == 0:
            g = torch.Generator().manual_seed(5)
            ref = torch.view_as_complex(torch.randn(nframes, n, 2, generator=g))
            srv = torch.view_as_complex(torch.randn(nframes, n, 2, generator=g))

        def fake_pipeline(r, s):            # stand-in for FramePipeline.run_device: a "map" per frame
            return (r * s.conj()).reshape(r.shape[0], 4, n // 4)[:, :, :3].contiguous()

        maps, idx = D.process_stream(fake_pipeline, ref, srv, nframes, n, src=0, mode=mode)
        bref = ref[0].clone() if rank == 0 else torch.zeros(n, dtype=torch.complex64)
        D.broadcast_reference(bref, src=0)
        if rank == 0:
            want = fake_pipeline(ref, srv)
            ok = torch.equal(maps, want) and torch.equal(bref, ref[0])
            q.put(("rank0", bool(ok), idx.tolist()))
        else:
            q.put(("rank1", maps is None, (idx.tolist(), float(bref.abs().sum()) > 0)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["interleaved", "contiguous"])
def test_scatter_process_gather_world2_gloo(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    nframes, n = 7, 64
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nframes, n, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(2):
        tag, ok, extra = q.get(timeout=120)
        results[tag] = (ok, extra)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results["rank0"][0], "gathered maps differ from single-process result"
    assert results["rank0"][1] == D.shard_indices(nframes, 0, 2, mode).tolist()
    assert results["rank1"][0]
    assert results["rank1"][1][0] == D.shard_indices(nframes, 1, 2, mode).tolist()
    assert results["rank1"][1][1]


def _stream_worker(rank, world, port, nframes, n, chunk, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        ref = torch.view_as_complex(torch.randn(nframes, n, 2, generator=g))      # every rank can rebuild the stream to check
        srv = torch.view_as_complex(torch.randn(nframes, n, 2, generator=g))
        got = []

        def source(ids):
            assert rank == 0
            return ref[ids].contiguous(), srv[ids].contiguous()

        def process(rr, ss, first):
            got.append((first, rr.clone(), ss.clone()))

        cnt = D.stream_frames(source, process, nframes, chunk, n, rank, world, torch.device("cpu"), src=0)
        ids = D.shard_indices(nframes, rank, world)
        ok = cnt == len(ids) and sum(g_[1].shape[0] for g_ in got) == len(ids)
        for first, rr, ss in got:
            want = ids[first:first + rr.shape[0]]
            ok = ok and torch.equal(rr, ref[want]) and torch.equal(ss, srv[want])
        q.put((rank, bool(ok), [g_[0] for g_ in got]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nframes,chunk", [(11, 2), (8, 4), (3, 5)])
def test_stream_frames_world2_gloo(nframes, chunk):
    """Config-3 staging plumbing: the ingest rank hands every rank its frames f % world chunk by chunk, in order."""
    world, n = 2, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stream_worker, args=(r, world, port, nframes, n, chunk, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    for rank, ok, firsts in res:
        assert ok, (rank, firsts)
        assert firsts == sorted(firsts)
