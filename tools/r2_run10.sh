#!/bin/bash
# 2 GPUs: the multi-GPU tests, rt1w_main --gpus 2, the bench at N = 2 as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -rP > gpurun_out/r2j_pytest_multi.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2j_pytest_multi.log | grep -v "^make\|^---"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2j_bench2.json 2> gpurun_out/r2j_bench2.err ) 2>&1 | tail -3; echo "bench exit $?"; tail -5 gpurun_out/r2j_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench2.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'C1 Mpaths/s',round(d['value'],1),'ms', round(d['ms_per_step'],2), 'e2e',round(d['e2e']['value'],1),'roofline',round(d['roofline']['frac'],3))
for c in d['configs']: print(c['name'], round(c['mpaths_per_s'],1),'Mpaths/s', round(c['mrays_per_s'],1),'Mrays/s')
for c in d['strong_scaling']: print(c['name'], c['total_spp'], round(c['ms_per_step'],1),'ms', round(c['mpaths_per_s'],1))
PY
