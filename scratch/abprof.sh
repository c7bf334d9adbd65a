#!/bin/bash
# bench + one full ncu capture of a steady-state wave for the product library and every variant
for lib in raytracing-1w_b200/_build/librt1w.so raytracing-1w_b200/_build/variant_*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so)
  RT1W_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', 'Mpaths/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1))"
  RT1W_LIB=$PWD/$lib ncu --set full --clock-control none --import-source on -k regex:"k_wave" -s 5 -c 1 -o gpurun_out/ab_$name -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ab_$name.log 2>&1
done
