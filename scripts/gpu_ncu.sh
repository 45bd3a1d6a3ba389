# ncu launch list + full capture of the dominant kernel (one GPU, short bench)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --rows 20000000 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_fixed_tiles} -s ${SKIP:-21} -c ${COUNT:-7} -f -o gpurun_out/prof_${TAG:-tiles} $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/ncu2.log
