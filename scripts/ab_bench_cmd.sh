python bench.py --steps 10 --warmup 3 --no-scans --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_bench_$V.json 2>> gpurun_out/ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/ab_bench_$V.json'))
print('$V', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), [round(c['ms'],3) for c in d['roofline']['per_column']])
PY
