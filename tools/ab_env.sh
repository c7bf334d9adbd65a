#!/bin/bash
# A/B of a runtime switch: bench + scene throughput with and without `VAR=VALUE` (e.g. tools/ab_env.sh RT1W_FACE_GROUPS=0)
SW=${1:-RT1W_FACE_GROUPS=0}
for env in "RT1W_NOOP=1" "$SW"; do
  echo "== $env"
  env $env python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  bench Mpaths/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'rays/path',round(d['rays_per_path'],4))"
  env $env timeout 120 python tools/scene_perf.py cornel_box:100 cornel_smoke:64 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['image'], 'spp', d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'rays/path', d['rays_per_path'])"
done
