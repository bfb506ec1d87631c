# quick A/B sweep of the lag-correlation geometry overrides (run on the GPU box)
for g in 8 4 2; do
  echo "== G=$g"; PRC_TUNE_G=$g python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),{n:k[n]['avg_us'] for n in k})"
done
