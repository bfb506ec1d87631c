#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/fir_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/fir_tests.log
PRC_FIR_DEBUG=1 timeout 120 python -c "
import numpy as np
from passiveradar_b200 import LS_Filter
from passiveradar_b200.synth import make_frame
ref,srv=make_frame(2**20,seed=3)[:2]
for _ in range(3): LS_Filter(ref,srv,300,1.0,10)
" 2>&1 | tail -2
timeout 300 python bench.py --steps 6 --warmup 3 > gpurun_out/fir_bench.log 2>&1; tail -1 gpurun_out/fir_bench.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:v['avg_us'] for k,v in d['kernels'].items()})"
