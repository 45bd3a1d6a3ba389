# per-kernel durations of the flat path (ncu launch lists; cold-cache, serialised)
mkdir -p gpurun_out
for C in dict_nulls plain_nulls; do for PG in 65536 8192; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_flat|k_decode_fixed|k_fixed_tiles' -s 9 -c 8 --csv --log-file gpurun_out/flat_${C}_${PG}.csv python scripts/bench_foreign.py 40000000 $PG $C > /dev/null 2>&1
echo "$C $PG"; python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/flat_${C}_${PG}.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hd=rows[h]; kn=hd.index('Kernel Name'); mv=hd.index('Metric Value')
for r in rows[h+1:]:
    if len(r)>mv: print('  ', r[kn][:60], r[mv])
PY
done; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_flat|k_decode_fixed|k_fixed_tiles' -s 27 -c 8 --csv --log-file gpurun_out/flat_opt.csv python scripts/bench_optional.py 40000000 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/flat_opt.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hd=rows[h]; kn=hd.index('Kernel Name'); mv=hd.index('Metric Value')
for r in rows[h+1:]:
    if len(r)>mv: print('  ', r[kn][:60], r[mv])
PY
