# quick GPU check: parity tests + smoke + one bench line (small scans unless BENCH_ARGS says otherwise)
mkdir -p gpurun_out
(free -g; nproc; df -h /dev/shm /tmp | cat; numactl -H 2>/dev/null | head -8; nvidia-smi topo -m 2>/dev/null | head -20; lscpu | grep -i "numa\|model name\|socket") > gpurun_out/box.txt 2>&1
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -12 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
r=d['roofline']
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',r['frac'],'kernel_ms',r['kernel_ms_per_step'],'launches',d['gpu_launches'])
for c in r['per_column']: print(c)
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print(d.get('cpu_baseline'))
for k in ('strings','regex','chunk_index'):
    if k in d:
        x=dict(d[k]); print(k, json.dumps(x)[:1800])
PY
