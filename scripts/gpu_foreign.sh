# foreign-writer pages: full GPU suite + scripts/bench_foreign.py at two sizes
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for R in 10000000 40000000; do
python scripts/bench_foreign.py $R > gpurun_out/foreign_${TAG:-a}_$R.json 2> gpurun_out/foreign.err; echo "foreign rc=$?"; tail -3 gpurun_out/foreign.err
python - <<PY
import json
d=json.load(open('gpurun_out/foreign_${TAG:-a}_$R.json'))
for r in d['results']:
    print($R, r['page_bytes'], r['column'], 'pages', r['pages'], 'ms', round(r['ms'],3), 'tiles', round(r['tiles_ms'],3), 'general', round(r['general_ms'],3), 'GB/s', round(r['in_plus_out_GBps']))
PY
done
python scripts/bench_optional.py 40000000 > gpurun_out/optional_${TAG:-a}.json 2> gpurun_out/optional.err; echo "optional rc=$?"; tail -2 gpurun_out/optional.err
python - <<PY
import json
d=json.load(open('gpurun_out/optional_${TAG:-a}.json'))
for r in d['results']: print(r['column'], 'ms', round(r['ms'],3), 'tiles', round(r['tiles_ms'],3), 'general', round(r['general_ms'],3), 'GB/s', round(r['in_plus_out_GBps']), 'frac', round(r['frac'],3))
PY
