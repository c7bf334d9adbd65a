#!/bin/bash
# round-end rehearsal on one GPU: smoke, GPU tests, both bench arms as the driver runs them, then the ncu captures for profiles/
TAG=${1:-final}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -4
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 600 gpurun_out/${TAG}_bench_reference.json
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { echo bench failed; tail -5 gpurun_out/${TAG}_bench.err; }
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('Mpaths/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'roofline',round(d['roofline']['frac'],3),'cpu',d['cpu_baseline'])"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_wave" -s 5 -c 3 -f -o gpurun_out/${TAG}_wave $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
