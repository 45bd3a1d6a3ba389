# A/B on one box: one plan over all columns (default) vs one plan per column (PQG_BENCH_PER_COLUMN=1)
for v in 0 1 0 1; do
  echo "== PQG_BENCH_PER_COLUMN=$v"
  PQG_BENCH_PER_COLUMN=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --regex-rows 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms',round(r['kernel_ms_per_step'],4),'tile launches',r['launches_per_step'],'all launches',d['gpu_launches'],'dict',round(r['dict_prepare_ms_per_step'],4),'general',round(r['general_kernel_ms_per_step'],4),'frac',round(r['frac'],4))"
done
