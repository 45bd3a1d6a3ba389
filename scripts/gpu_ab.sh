# A/B of build variants: duckdb-parquet-parser_b200/variants/libpqg_<X>.so against the in-tree libpqg.so (= A); CMD is run per variant
mkdir -p gpurun_out
PKG=duckdb-parquet-parser_b200
cp $PKG/libpqg.so /tmp/libpqg_A.so
for V in A ${VARIANTS:-B C D}; do
  if [ $V != A ]; then cp $PKG/variants/libpqg_$V.so $PKG/libpqg.so; else cp /tmp/libpqg_A.so $PKG/libpqg.so; fi
  echo "== variant $V"
  V=$V bash -c "$CMD"
done
cp /tmp/libpqg_A.so $PKG/libpqg.so
