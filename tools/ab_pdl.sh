for i in 1 2; do
  for v in 0 1; do
    BM_NO_PDL=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-modes 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('NO_PDL=$v sift',round(d['value']),round(d['e2e']['value']),'orb',round(d['orb']['value']),round(d['orb']['e2e']['value']),'chain',d['roofline']['ms_per_frame'])
"
  done
done
