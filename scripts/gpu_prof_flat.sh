# flat path: ncu source-level captures of the scan and emit kernels (64 KB pyarrow pages with nulls, dictionary column)
mkdir -p gpurun_out
C=${COL:-dict_nulls}
ncu --set full --clock-control none --import-source on -k regex:k_flat_emit -s 2 -c 1 -f -o gpurun_out/prof_flat_emit_${TAG:-a} python scripts/bench_foreign.py 40000000 65536 $C > gpurun_out/ncu_flat_emit.log 2>&1; echo "ncu emit rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_flat_scan -s 2 -c 1 -f -o gpurun_out/prof_flat_scan_${TAG:-a} python scripts/bench_foreign.py 40000000 65536 $C > gpurun_out/ncu_flat_scan.log 2>&1; echo "ncu scan rc=$?"
