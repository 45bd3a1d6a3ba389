# foreign-writer pages: the foreign / page parity tests + scripts/bench_foreign.py at 40 M rows
mkdir -p gpurun_out
python -m pytest tests/test_gpu_foreign.py tests/test_gpu_pages.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_f.log
R=${ROWS:-40000000}
python scripts/bench_foreign.py $R > gpurun_out/foreign_${TAG:-a}_$R.json 2> gpurun_out/foreign.err; echo "foreign rc=$?"; tail -3 gpurun_out/foreign.err
python - <<PY
import json
d=json.load(open('gpurun_out/foreign_${TAG:-a}_$R.json'))
for r in d['results']:
    print($R, r['page_bytes'], r['column'], 'pages', r['pages'], 'ms', round(r['ms'],3), 'tiles', round(r['tiles_ms'],3), 'general', round(r['general_ms'],3), 'GB/s', round(r['in_plus_out_GBps']))
PY
