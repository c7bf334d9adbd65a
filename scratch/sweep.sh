#!/bin/bash
# bench every tuning variant library in _build/
for lib in raytracing-1w_b200/_build/librt1w.so raytracing-1w_b200/_build/variant_*.so; do
  [ -f "$lib" ] || continue
  RT1W_LIB=$PWD/$lib python bench.py --steps 3 --warmup 2 --no-cpu-baseline "$@" 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib'.split('/')[-1], 'Mpaths/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2), {k:v['ms'] for k,v in d['kernels'].items()})"
done
