# A/B of string-kernel build variants: duckdb-parquet-parser_b200/variants/libpqg_<X>.so against the in-tree libpqg.so (= A)
mkdir -p gpurun_out
PKG=duckdb-parquet-parser_b200
cp $PKG/libpqg.so /tmp/libpqg_A.so
for V in A ${VARIANTS:-B C}; do
  if [ $V != A ]; then cp $PKG/variants/libpqg_$V.so $PKG/libpqg.so; else cp /tmp/libpqg_A.so $PKG/libpqg.so; fi
  for W in cfg3 cfg4 cfg1; do
    python scripts/bench_strings.py ${ROWS:-40000000} $W > gpurun_out/var_${V}_$W.json 2>> gpurun_out/var.err
  done
  python - <<PY
import json
for w in ('cfg3','cfg4','cfg1'):
    r=json.load(open('gpurun_out/var_${V}_%s.json' % w))['results'][0]
    print('$V', w, 'dict',round(r['dict_prepare_ms'],3),'size',round(r['size_pass_ms'],3),'copy',round(r['copy_pass_ms'],3),'frac',round(r['frac_of_hbm_peak'],3))
PY
done
cp /tmp/libpqg_A.so $PKG/libpqg.so
