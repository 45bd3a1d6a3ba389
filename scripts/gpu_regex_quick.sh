# regex scan only: parity tests of the scan paths + numbers (no ncu)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py tests/test_gpu_foreign.py -m gpu -q > gpurun_out/pytest_rx.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_rx.log
python scripts/prof_regex.py ${ROWS:-100000000} > gpurun_out/regex_${TAG:-a}.json 2> gpurun_out/regex.err; echo "regex rc=$?"; tail -3 gpurun_out/regex.err
cat gpurun_out/regex_${TAG:-a}.json
