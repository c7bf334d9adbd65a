#!/bin/bash
# 8 GPUs: the multi-GPU tests (N = 2, 4, 8 inside one process, one process per GPU, rt1w_main --gpus), then the bench at N = 8 and 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -rP > gpurun_out/r2l_pytest_multi.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2l_pytest_multi.log | grep -v "^make\|^---"
for n in 8 4; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2l_bench$n.json 2> gpurun_out/r2l_bench$n.err ) 2>&1 | tail -3; echo "bench $n exit $?"; tail -2 gpurun_out/r2l_bench$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2l_bench$n.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'C1 Mpaths/s',round(d['value'],1),'ms', round(d['ms_per_step'],2), 'e2e',round(d['e2e']['value'],1),'roofline',round(d['roofline']['frac'],3))
for c in d['configs']: print(c['name'], round(c['mpaths_per_s'],1),'Mpaths/s', round(c['mrays_per_s'],1),'Mrays/s')
for c in d['strong_scaling']: print(c['name'], c['total_spp'], round(c['ms_per_step'],1),'ms', round(c['mpaths_per_s'],1))
PY
done
( time ./raytracing-1w_b200/_build/rt1w_main cornel_box --gpus 8 > gpurun_out/r2l_cornell_8gpu.ppm 2> gpurun_out/r2l_main8.err ) 2>&1 | tail -3; head -c 40 gpurun_out/r2l_cornell_8gpu.ppm | head -3; rm -f gpurun_out/r2l_cornell_8gpu.ppm
