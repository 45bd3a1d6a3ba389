# string kernels: numbers + ncu source-level captures (k_str_pages on the cfg3 and cfg4 shapes)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_scan.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_scan.log
python scripts/bench_strings.py ${ROWS:-40000000} > gpurun_out/strings_${TAG:-a}.json 2> gpurun_out/strings.err; echo "strings rc=$?"; tail -3 gpurun_out/strings.err
python - <<PY
import json
d=json.load(open('gpurun_out/strings_${TAG:-a}.json'))
for r in d['results']:
    print(r['workload'][:40], 'rows',r['rows'],'pages',r['pages'],'dict',round(r['dict_prepare_ms'],3),'size',round(r['size_pass_ms'],3),'copy',round(r['copy_pass_ms'],3),'in+out GB/s',round(r['in_plus_out_GBps']),'frac',round(r['frac_of_hbm_peak'],3),'chunk_index_ms',round(r['chunk_index_ms'],3))
PY
for W in cfg3 cfg4; do
  ncu --set full --clock-control none --import-source on -k regex:k_str_pages -s 2 -c 2 -f -o gpurun_out/prof_str_${W}_${TAG:-a} python scripts/bench_strings.py ${ROWS:-40000000} $W > gpurun_out/ncu_str_${W}.log 2>&1
  echo "ncu $W rc=$?"
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --strings-rows 0 > gpurun_out/bench_ci.json 2> gpurun_out/bench_ci.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ci.json'))
x=d['chunk_index']; print({k:x[k] for k in ('value','ms_per_step_wall','kernel_ms','serial_chain_ms','total_chunks')})
x=d['regex']; print('regex', x['value'], x['kernel_ms'], x['neg_regex']['kernel_ms'], x['frac_of_hbm_peak'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan|chain|str_|dict_|narrow|chunk_row' -c 120 --csv --log-file gpurun_out/launches_ci.csv python bench.py --rows 10000000 --steps 2 --warmup 3 --no-cpu-baseline --strings-rows 0 --regex-rows 0 --e2e-steps 1 --scan-steps 1 > gpurun_out/ncu_ci.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_ci.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hd=rows[h]; kn=hd.index('Kernel Name'); mv=hd.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)<=mv: continue
    n=r[kn].split('(')[0][-50:]
    agg.setdefault(n,[]).append(float(r[mv].replace(',','')))
for n,v in agg.items():
    print(n, len(v), 'last(us)', v[-1]/1000.0)
PY
