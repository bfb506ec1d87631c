# quick A/B sweep of kernel variants (run on the GPU box)
for t in 0 1 2; do
  echo "== PRC_TILE=$t"; PRC_TILE=$t python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),{n:k[n]['avg_us'] for n in k})"
done
