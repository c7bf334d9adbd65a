#!/bin/bash
# tests + bench + ncu (launch list, one full capture of a mid-render wave with sources)
TAG=${1:-x}
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { echo bench failed; tail -5 gpurun_out/${TAG}_bench.err; exit 1; }
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('Mpaths/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), {k:v['ms'] for k,v in d['kernels'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_wave|k_finish" -s ${2:-5} -c 3 -o gpurun_out/${TAG}_wave $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
