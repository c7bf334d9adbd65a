#!/bin/bash
# profiles/r01_* from the captures of one tools/final_check.sh (or prof_bench.sh) run: tools/profiles_refresh.sh TAG
TAG=${1:?tag of the gpurun_out files}
REP=gpurun_out/${TAG}_wave.ncu-rep
tail -n 1 gpurun_out/${TAG}_bench.json > profiles/r01_bench_line.json
[ -f gpurun_out/${TAG}_bench_reference.json ] && tail -n 1 gpurun_out/${TAG}_bench_reference.json > profiles/r01_bench_line_reference.json
cp gpurun_out/${TAG}_launches.csv profiles/r01_launches.csv
python tools/ncu_summary.py $REP > profiles/r01_wave_summary.txt
python tools/ncu_regions.py $REP k_wave > profiles/r01_wave_regions.txt
python tools/ncu_lines.py $REP k_wave > profiles/r01_wave_lines.txt
ncu -i $REP --page details --csv --launch-skip 0 --launch-count 1 > profiles/r01_wave_details.csv
ncu -i $REP --page raw --csv --launch-skip 0 --launch-count 1 | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h, u, r = rows[0], rows[1], rows[2]
want = [l.strip() for l in open('tools/profile_metrics.txt') if l.strip()]
for k in want:
    if k in h: print(k, r[h.index(k)], u[h.index(k)])
" > profiles/r01_wave_metrics.txt
