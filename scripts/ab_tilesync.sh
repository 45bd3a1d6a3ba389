# A/B on one box: last-warp refill (default) vs CTA-wide barrier per tile (PQG_TILE_SYNC=1)
for v in 0 1 0 1; do
  echo "== PQG_TILE_SYNC=$v"
  PQG_TILE_SYNC=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms',round(r['kernel_ms_per_step'],4),'regex_ms',round(d['regex']['kernel_ms'],4), 'neg', round(d['regex']['neg_regex']['kernel_ms'],4))
print(' '.join('%s=%.4f'%(c['column'],c['ms']) for c in r['per_column']))"
done
