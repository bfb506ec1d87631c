"""Timing of the batched frame path at config 2 (for A/B builds: PRC_LIBRARY=... python scripts/fft/time_only.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from passiveradar_b200 import _lib, synth
from passiveradar_b200.frames import FramePipeline

n, F, R, B = 2 ** 20, 256, 300, 125
dev = torch.device("cuda", 0)
fr = [synth.make_frame(n, "P1", i) for i in range(4)]
ref_d = torch.from_numpy(np.stack([fr[i % 4][0] for i in range(B)])).to(dev)
srv_d = torch.from_numpy(np.stack([fr[i % 4][1] for i in range(B)])).to(dev)
maps = torch.empty((B, F, R + 1), dtype=torch.complex64, device=dev)
if len(sys.argv) > 1:
    for kv in sys.argv[1:]:
        k, v = kv.split("=")
        _lib.set_option(k, int(v))
for batch, slots in ((25, 5),):
    pipe = FramePipeline(n, R, F, batch=batch, nslots=slots)
    for _ in range(3):
        pipe.run_device(ref_d, srv_d, maps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        pipe.run_device(ref_d, srv_d, maps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{os.environ.get('PRC_LIBRARY', 'default')}: batch={batch} slots={slots}: {B * reps / ms * 1e3:.0f} frames/s", flush=True)
_lib.profile_reset(); _lib.profile(True)
pipe = FramePipeline(n, R, F, batch=16, nslots=1)
pipe.run_device(ref_d, srv_d, maps); torch.cuda.synchronize()
prof = _lib.profile_read(); _lib.profile(False)
print("   per frame: " + ", ".join(f"{k} {1e3 * v[0] / B:.2f}us" for k, v in prof.items() if v[1]), flush=True)
