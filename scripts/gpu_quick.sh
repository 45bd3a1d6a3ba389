# quick GPU check: parity tests + smoke + one bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
r=d['roofline']
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',r['frac'],'kernel_ms',r['kernel_ms_per_step'],'launches',d['gpu_launches'])
for c in r['per_column']: print(c)
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print(d.get('cpu_baseline'))
PY
