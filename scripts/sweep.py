#!/usr/bin/env python
"""BASELINE config 5: sweep CPI length x Doppler bins, frames/s and achieved HBM GB/s (device resident).

    python scripts/sweep.py [--gpus N]      (N > 1: launch under torchrun; weak scaling, max over ranks)

GB/s = bytes_frame(N, F, R) * frames/s with bytes_frame = 2*8*N + 8*F*(R+1);
roofline fraction is against MEASURED_PEAKS.json's copy bandwidth.  R = 300, LS filterLen = 300.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from passiveradar_b200 import synth
    from passiveradar_b200.frames import FramePipeline
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = 6571.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    R = 300
    rows = []
    for logn in (18, 19, 20, 21, 22):
        n = 2 ** logn
        B = max(4, min(16, (2 ** 24) // n))          # frames per step: >= 268 MB of input where possible
        ref0, srv0 = synth.make_frame(n, "P1", rank)
        ref_d = torch.from_numpy(np.stack([ref0] * B)).to(dev)
        srv_d = torch.from_numpy(np.stack([srv0] * B)).to(dev)
        for F in (64, 256, 1024):
            pipe = FramePipeline(n, R, F, filter_len=R, window=("kaiser", 5.0), device=local, nslots=8)
            maps = torch.empty((B, F, R + 1), dtype=torch.complex64, device=dev)
            for _ in range(3):
                pipe.run_device(ref_d, srv_d, maps)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            steps = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                pipe.run_device(ref_d, srv_d, maps)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            fps = B * steps * world / (ms * 1e-3)
            gbs = (2 * 8 * n + 8 * F * (R + 1)) * fps / 1e9
            rows.append((n, F, world, fps, gbs, gbs / (peak * world)))
            if rank == 0:
                print(f"N=2^{logn} F={F:5d} R={R} gpus={world}: {fps:10.1f} frames/s  {gbs:8.1f} GB/s  {100 * gbs / (peak * world):5.2f} % of measured HBM peak", flush=True)
            del pipe, maps
        del ref_d, srv_d
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"sweep": [dict(n=r[0], F=r[1], gpus=r[2], fps=r[3], GBps=r[4], frac=r[5]) for r in rows]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
