# OPTIONAL fixed-width tile kernel: numbers + ncu source-level captures (PLAIN and dictionary bw 16 columns)
mkdir -p gpurun_out
python scripts/bench_optional.py ${ROWS:-40000000} > gpurun_out/optional_${TAG:-a}.json 2> gpurun_out/optional.err; echo "optional rc=$?"; tail -3 gpurun_out/optional.err
python - <<PY
import json
for r in json.load(open('gpurun_out/optional_${TAG:-a}.json'))['results']: print(r['column'], 'ms', round(r['ms'],3), 'tiles', round(r['tiles_ms'],3), 'general', round(r['general_ms'],3), 'GB/s', round(r['in_plus_out_GBps']), 'frac', round(r['frac'],3))
PY
ncu --set full --clock-control none --import-source on -k regex:k_fixed_tiles -s 2 -c 1 -f -o gpurun_out/prof_opt_plain_${TAG:-a} python scripts/bench_optional.py ${ROWS:-40000000} > gpurun_out/ncu_opt_plain.log 2>&1; echo "ncu plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fixed_tiles -s 12 -c 1 -f -o gpurun_out/prof_opt_d16_${TAG:-a} python scripts/bench_optional.py ${ROWS:-40000000} > gpurun_out/ncu_opt_d16.log 2>&1; echo "ncu d16 rc=$?"
