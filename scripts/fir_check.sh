#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/fir_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/fir_tests.log
timeout 300 python bench.py --steps 6 --warmup 3 > gpurun_out/fir_bench.log 2>&1; tail -1 gpurun_out/fir_bench.log > gpurun_out/bench_r01_tma.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_tma.json')); print(d['value'], d['e2e']['value'], {k:v['avg_us'] for k,v in d['kernels'].items()}); print(d['cpu_baseline']['value'])"
