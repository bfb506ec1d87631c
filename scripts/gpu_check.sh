#!/bin/bash
# One-call GPU verification used during development (run through gpurun from the repo root):
#   smoke(), the GPU parity tests, the default bench line, the ncu launch list of a short bench run and one
#   `--set full` capture of the two transform kernels.  Results under gpurun_out/ (copy what is to be judged to profiles/).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e_dropin']['value'], d['latency_us'], d['gpu_launches'], d['clocks'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['frame_frac'], d['cpu_baseline']['value'])"
# launch list of the same command (short): kernel SHARES must agree with the bench line's `kernels`
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --frames-per-step 250 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"lscorr_fft|caf_fft" -s 2 -c 2 -o gpurun_out/r02_fft_final \
    python scripts/fft/prof_one.py 16 2 > gpurun_out/ncu_final.log 2>&1; echo "ncu full rc=$?"
