#!/bin/bash
# usage: envsweep.sh VAR v1 v2 ...   (bench only)
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$VAR=$v', 'Mpaths/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'], {k:v['ms'] for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'])"
done
