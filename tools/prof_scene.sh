#!/bin/bash
# one full ncu capture of a steady-state wave for each scene given as name:spp:skip
for spec in "$@"; do
  IFS=: read name spp skip <<< "$spec"
  python tools/scene_perf.py $name:$spp | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['scene'], d['image'], 'spp', d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'waves', d['waves'], 'nan', d['nan_pixels'])"
  ncu --set full --clock-control none --import-source on -k regex:"k_wave" -s ${skip:-8} -c 1 -f -o gpurun_out/scene_$name python tools/scene_perf.py $name:$spp > gpurun_out/scene_$name.log 2>&1
done
