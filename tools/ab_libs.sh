#!/bin/bash
# bench + Cornell scene throughput for the product library and every variant, interleaved twice (box-to-box and run-to-run noise)
for rep in 1 2; do
for lib in raytracing-1w_b200/_build/librt1w.so raytracing-1w_b200/_build/variant_*.so; do
  [ -f "$lib" ] || continue
  RT1W_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib'.split('/')[-1], 'Mpaths/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'rays/path',round(d['rays_per_path'],4))"
done
done
