# strings only: parity tests of the string paths + numbers (no ncu)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_pages.py tests/test_gpu_scan.py tests/test_gpu_foreign.py tests/test_known_answer.py -m gpu -q > gpurun_out/pytest_str.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_str.log
python scripts/bench_strings.py ${ROWS:-40000000} > gpurun_out/strings_${TAG:-a}.json 2> gpurun_out/strings.err; echo "strings rc=$?"; tail -3 gpurun_out/strings.err
python - <<PY
import json
d=json.load(open('gpurun_out/strings_${TAG:-a}.json'))
for r in d['results']:
    print(r['workload'][:40], 'rows',r['rows'],'pages',r['pages'],'dict',round(r['dict_prepare_ms'],3),'size',round(r['size_pass_ms'],3),'copy',round(r['copy_pass_ms'],3),'in+out GB/s',round(r['in_plus_out_GBps']),'frac',round(r['frac_of_hbm_peak'],3),'chunk_index_ms',round(r['chunk_index_ms'],3))
PY
