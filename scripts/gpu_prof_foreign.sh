# foreign pages with nulls: ncu source-level captures of the big-page kernel (64 KB pages) and the general kernel (8 KB pages)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_big_pages -s 2 -c 1 -f -o gpurun_out/prof_big_dn_${TAG:-a} python scripts/bench_foreign.py 40000000 65536 dict_nulls > gpurun_out/ncu_big_dn.log 2>&1; echo "ncu big dict_nulls rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_big_pages -s 2 -c 1 -f -o gpurun_out/prof_big_pn_${TAG:-a} python scripts/bench_foreign.py 40000000 65536 plain_nulls > gpurun_out/ncu_big_pn.log 2>&1; echo "ncu big plain_nulls rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_decode_fixed -s 2 -c 1 -f -o gpurun_out/prof_gen_dn_${TAG:-a} python scripts/bench_foreign.py 40000000 8192 dict_nulls > gpurun_out/ncu_gen_dn.log 2>&1; echo "ncu general dict_nulls rc=$?"
