#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_wide_bvh.py -m gpu -q --tb=short -rP -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2h_pytest.log | grep -v "^make\|^---"
for flags in 36 40; do
  echo "== flags $flags"
  RT1W_FLAGS=$flags timeout 600 python tools/scene_perf.py one_weekend:32 final_scene:32 stress:8 2>gpurun_out/r2h_perf.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'wide', d['wide_nodes'], d['wide_depth'], d['wide_children'], 'build', d['build_ms'])"
done
RT1W_NO_WARMUP=1 RT1W_FLAGS=40 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 1 -c 1 -f -o gpurun_out/r2h_stress_wide python tools/scene_perf.py stress:8 > gpurun_out/r2h_ncu_stress.log 2>&1
