#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -q --tb=short -rP --durations=6 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2k_pytest.log | grep -v "^make\|^---"
( time python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err ) 2>&1 | tail -3; echo "bench exit $?"; tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print('C1 Mpaths/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'roofline',round(d['roofline']['frac'],3),'cpu',d['cpu_baseline']['value'])
for c in d['configs']: print(c['name'], round(c['mpaths_per_s'],1),'Mpaths/s', round(c['mrays_per_s'],1),'Mrays/s roofline', round(c['roofline']['frac'],3), 'cpu', c.get('cpu_baseline',{}).get('value'), 'x', c.get('gpu_over_cpu'), c['scene_build'])
for c in d['strong_scaling']: print(c['name'], c['total_spp'], round(c['ms_per_step'],1),'ms', round(c['mpaths_per_s'],1))
PY
