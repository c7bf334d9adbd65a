#!/bin/bash
# profiles/<TAG>_* from the captures of tools/profiles_capture.sh (run here, after the gpurun call has merged gpurun_out/): tools/profiles_refresh.sh r02
TAG=${1:?tag of the gpurun_out files}
OUT=${2:-profiles} # (tools/profiles_capture.sh digests on the GPU box into gpurun_out/profiles_<TAG>: the reports are too big to travel)
mkdir -p $OUT
tail -n 1 gpurun_out/${TAG}_bench.json > $OUT/${TAG}_bench_line.json
[ -f gpurun_out/${TAG}_bench_reference.json ] && tail -n 1 gpurun_out/${TAG}_bench_reference.json > $OUT/${TAG}_bench_line_reference.json
cp gpurun_out/${TAG}_launches.csv $OUT/${TAG}_launches.csv
python tools/ncu_flops.py ${TAG} $OUT > /dev/null
for c in C1 C2 C2w C3 C5; do
  REP=gpurun_out/${TAG}_${c}_wave.ncu-rep
  [ -f $REP ] || continue
  python tools/ncu_summary.py $REP > $OUT/${TAG}_${c}_summary.txt
  python tools/ncu_regions.py $REP k_wave | head -45 > $OUT/${TAG}_${c}_regions.txt
  ncu -i $REP --page raw --csv --launch-skip 0 --launch-count 1 | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h, u, r = rows[0], rows[1], rows[2]
want = [l.strip() for l in open('tools/profile_metrics.txt') if l.strip()]
for k in want:
    if k in h: print(k, r[h.index(k)], u[h.index(k)])
" > $OUT/${TAG}_${c}_metrics.txt
done
python tools/ncu_lines.py gpurun_out/${TAG}_C1_wave.ncu-rep k_wave > $OUT/${TAG}_C1_lines.txt 2>/dev/null
ls -la $OUT | grep ${TAG}
