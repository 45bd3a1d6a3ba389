# gpurun_out/<tag>_* (scripts/gpu_round2_final.sh) -> the tracked summaries under profiles/ ; needs the SAME libpqg.so build as the captures
T=${TAG:-r02_v3}; G=gpurun_out
python scripts/summarize_profiles.py $T $G/${T}_prof_tiles.ncu-rep $G/${T}_launches.csv $G/${T}_prof_regex.ncu-rep
python - <<PY
import sys
sys.argv=['x']; sys.path.insert(0,'scripts')
import summarize_profiles as sp
T='$T'; G='$G'
for src,dst in ((f'{G}/{T}_prof_str_cfg3.ncu-rep',f'profiles/{T}_ncu_full_k_str_pages_cfg3.csv'),(f'{G}/{T}_prof_str_cfg4.ncu-rep',f'profiles/{T}_ncu_full_k_str_pages_cfg4.csv'),
                (f'{G}/{T}_prof_flat_emit.ncu-rep',f'profiles/{T}_ncu_full_k_flat_emit.csv'),(f'{G}/{T}_prof_flat_scan.ncu-rep',f'profiles/{T}_ncu_full_k_flat_scan.csv')):
    sp.summarize(src,dst)
PY
for f in bench bench_ref regex strings optional foreign foreign_10M ext; do [ -f $G/${T}_$f.json ] && cp $G/${T}_$f.json profiles/; done
[ -f $G/${T}_pytest_gpu.log ] && cp $G/${T}_pytest_gpu.log profiles/
python scripts/ncu_lines.py $G/${T}_prof_regex.ncu-rep k_regex_tiles pqg_scan 714288 > profiles/${T}_lines_k_regex_tiles.txt 2>&1; head -1 profiles/${T}_lines_k_regex_tiles.txt
python scripts/ncu_lines.py $G/${T}_prof_str_cfg3.ncu-rep k_str_pagesILb1ELb1 pqg_decode 78128 2 > profiles/${T}_lines_k_str_pagesILb1ELb1_cfg3.txt 2>&1; head -1 profiles/${T}_lines_k_str_pagesILb1ELb1_cfg3.txt
python scripts/ncu_lines.py $G/${T}_prof_str_cfg4.ncu-rep k_str_pagesILb1ELb0 pqg_decode 1428576 0 > profiles/${T}_lines_k_str_pagesILb1ELb0_cfg4.txt 2>&1; head -1 profiles/${T}_lines_k_str_pagesILb1ELb0_cfg4.txt
python scripts/ncu_lines.py $G/${T}_prof_flat_emit.ncu-rep k_flat_emitILi8ELb1 pqg_flat 39063 > profiles/${T}_lines_k_flat_emit.txt 2>&1; head -1 profiles/${T}_lines_k_flat_emit.txt
python scripts/ncu_lines.py $G/${T}_prof_flat_scan.ncu-rep k_flat_scanILi8 pqg_flat 2000 > profiles/${T}_lines_k_flat_scan.txt 2>&1; head -1 profiles/${T}_lines_k_flat_scan.txt
