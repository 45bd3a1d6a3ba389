# Full-size captures for profiles/: launch list of the bench, ncu --set full of the 4 k_fixed_tiles
# launches of one timed step (100 M rows x 7 columns in one plan; the 32 + 12 launches before them are
# the per-column detail pass and the warm-up), and of the regex tile kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --regex-rows 5000000"
$CMD > gpurun_out/plain_full.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_full.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fixed_tiles -s 44 -c 4 -f -o gpurun_out/prof_tiles_full $CMD > gpurun_out/ncu_t.log 2>&1
echo "tiles capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_regex_tiles -s 3 -c 1 -f -o gpurun_out/prof_regex_full $CMD > gpurun_out/ncu_r.log 2>&1
echo "regex capture rc=$?"
