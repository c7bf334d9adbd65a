#!/bin/bash
# A/B of build variants (python raytracing-1w_b200/build.py --variant NAME -D...): tools/ab_libs.sh "scene:spp ..." librt1w variant_a variant_b
SCENES=${1:-cornel_box:100}; shift
for lib in "$@"; do
  echo "== $lib"
  for i in 1 2; do
    RT1W_LIB=$PWD/raytracing-1w_b200/_build/$lib.so timeout 600 python tools/scene_perf.py $SCENES 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['image'], d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'])"
  done
done
