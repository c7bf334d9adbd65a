#!/bin/bash
# quick GPU check: parity tests + short bench summary
python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -4
python bench.py --steps ${1:-3} --warmup 2 --no-cpu-baseline $2 $3 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('Mpaths/s',round(d['value'],1),'Mrays/s',round(d['mrays_per_s'],1),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'])
print({k:v['ms'] for k,v in d['kernels'].items()}); print('roofline frac',d['roofline']['frac'])"
