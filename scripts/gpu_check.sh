#!/bin/bash
# One-call GPU verification used during development (run through gpurun from the repo root):
#   smoke(), the GPU parity tests, the default bench line and an ncu launch list of a short bench run.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
timeout 300 python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/final_bench.log > gpurun_out/bench_r01_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_final.json')); print(d['value'], d['e2e']['value'], d['steps'], d['warmup'], d['gpu_launches'], d['clocks'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('issued_frac'), d['cpu_baseline']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 1 --batch 4 --no-cpu-baseline > gpurun_out/b_ncu9.log 2>&1; echo "ncu rc=$?"
python scripts/rows_bench.py 2>&1 | tail -1 > gpurun_out/rows_final.json; cut -c1-400 gpurun_out/rows_final.json
python - <<'PY'
import time, numpy as np
import passiveradar_b200 as prb
from passiveradar_b200 import synth
ref, srv = synth.make_frame(2**21, "P1", 0)
for _ in range(2):
    t = time.time(); prb.NLMS_filter(ref, srv, 400, 0.05, 10); print("NLMS config 4 (2^21 samples, 410 taps):", round(time.time() - t, 4), "s")
PY
