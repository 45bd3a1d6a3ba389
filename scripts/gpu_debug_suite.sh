# the GPU suite once under the debug build (device-side index asserts, duckdb-parquet-parser_b200/variants/libpqg_debug.so)
mkdir -p gpurun_out
PKG=duckdb-parquet-parser_b200
cp $PKG/libpqg.so /tmp/libpqg_release.so && cp $PKG/variants/libpqg_debug.so $PKG/libpqg.so
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_debug.log 2>&1; echo "pytest(debug build) rc=$?"; tail -4 gpurun_out/pytest_gpu_debug.log; grep -c PQG_ASSERT gpurun_out/pytest_gpu_debug.log
cp /tmp/libpqg_release.so $PKG/libpqg.so
