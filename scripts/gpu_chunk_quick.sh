# chunk index only: scan / known-answer / multi tests + the bench's chunk_index object
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py tests/test_known_answer.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_ci.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ci.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --strings-rows 0 --regex-rows 0 --e2e-steps 1 > gpurun_out/bench_ci.json 2> gpurun_out/bench_ci.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ci.json'))
x=d['chunk_index']; print({k:x[k] for k in ('value','ms_per_step_wall','kernel_ms','serial_chain_ms','total_chunks')})
PY
