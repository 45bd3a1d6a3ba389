set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench1_ref.json 2> gpurun_out/bench1_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --rows 20000000 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --rows 20000000 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_fixed -s 21 -c 7 -o gpurun_out/prof_r1_fixed python bench.py --steps 2 --warmup 3 --rows 20000000 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cat gpurun_out/bench1.json | cut -c1-1500
