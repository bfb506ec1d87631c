#!/bin/bash
export PYTHONPATH=$PWD
for cv in 1 0; do
PRC_CARVEOUT=$cv timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('carveout $cv', round(d['value']), round(d['e2e']['value']), {k:v['avg_us'] for k,v in d['kernels'].items()})"
done
