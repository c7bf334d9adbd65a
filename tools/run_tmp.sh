#!/bin/bash
mkdir -p gpurun_out
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n > gpurun_out/s2_bench_n$n.json 2> gpurun_out/s2_bench_n$n.err || tail -5 gpurun_out/s2_bench_n$n.err
  tail -n 1 gpurun_out/s2_bench_n$n.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'C1', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
for c in d.get('configs', []): print('   ', c['name'], round(c['mpaths_per_s'], 1))
for c in d.get('strong_scaling', []): print('   strong', c['name'], round(c['ms_per_step'], 2), 'ms', round(c['mpaths_per_s'], 1))"
done
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
