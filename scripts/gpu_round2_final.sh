# Round-2 captures for profiles/: bench lines (both arms), ncu launch list of the bench, ncu --set full of the dominant kernel
# (the 4 k_fixed_tiles launches of one timed step), of the string kernels, the regex tile kernel and the flat block decode, and the side benches.
mkdir -p gpurun_out
T=${TAG:-r02}
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "reference arm rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
export PQG_BENCH_NO_AB=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --scan-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fixed_tiles -s 44 -c 4 -f -o gpurun_out/${T}_prof_tiles python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-scans > gpurun_out/ncu_t.log 2>&1; echo "tiles capture rc=$?"
for W in cfg3 cfg4; do
  ncu --set full --clock-control none --import-source on -k regex:k_str_pages -s 2 -c 2 -f -o gpurun_out/${T}_prof_str_$W python scripts/bench_strings.py 40000000 $W > gpurun_out/ncu_s_$W.log 2>&1; echo "strings capture $W rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:k_regex_tiles -s 2 -c 1 -f -o gpurun_out/${T}_prof_regex python scripts/prof_regex.py 20000000 > gpurun_out/ncu_r.log 2>&1; echo "regex capture rc=$?"
for K in emit scan; do
  ncu --set full --clock-control none --import-source on -k regex:k_flat_$K -s 2 -c 1 -f -o gpurun_out/${T}_prof_flat_$K python scripts/bench_foreign.py 40000000 65536 dict_nulls > gpurun_out/ncu_f_$K.log 2>&1; echo "flat $K capture rc=$?"
done
python scripts/prof_regex.py 100000000 > gpurun_out/${T}_regex.json 2>> gpurun_out/side.err
python scripts/bench_strings.py 40000000 > gpurun_out/${T}_strings.json 2>> gpurun_out/side.err; echo "strings rc=$?"
python scripts/bench_optional.py 40000000 > gpurun_out/${T}_optional.json 2>> gpurun_out/side.err; echo "optional rc=$?"
python scripts/bench_foreign.py 40000000 > gpurun_out/${T}_foreign.json 2>> gpurun_out/side.err; echo "foreign rc=$?"
python scripts/bench_foreign.py 10000000 > gpurun_out/${T}_foreign_10M.json 2>> gpurun_out/side.err; echo "foreign 10M rc=$?"
python scripts/bench_ext.py 10000000 > gpurun_out/${T}_ext.json 2>> gpurun_out/side.err; echo "ext rc=$?"
tail -3 gpurun_out/side.err
