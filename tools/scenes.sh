#!/bin/bash
# scene throughput for the product library and every variant
for lib in raytracing-1w_b200/_build/librt1w.so raytracing-1w_b200/_build/variant_*.so; do
  [ -f "$lib" ] || continue
  echo "== $(basename $lib)"
  RT1W_LIB=$PWD/$lib timeout 60 python tools/scene_perf.py "$@" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['image'], 'spp', d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'nan', d['nan_pixels'])"
done
