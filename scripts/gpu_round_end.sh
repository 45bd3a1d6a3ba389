# round-end capture: tests, smoke, bench (+reference arm), secondary measurements, ncu evidence
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python scripts/bench_strings.py 10000000 > gpurun_out/strings.json 2>/dev/null; echo "strings rc=$?"
python scripts/bench_optional.py 20000000 > gpurun_out/optional.json 2>/dev/null; echo "optional rc=$?"
python scripts/bench_foreign.py 10000000 > gpurun_out/foreign.json 2>/dev/null; echo "foreign rc=$?"
bash scripts/gpu_profile_final.sh
