#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_wide_bvh.py -m gpu -q --tb=short -rP -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2e_pytest.log | grep -v "^make\|^---"
for lib in librt1w variant_is8 variant_is4; do
for flags in 24 36 40; do
  echo "== $lib flags $flags"
  RT1W_LIB=$PWD/raytracing-1w_b200/_build/$lib.so RT1W_FLAGS=$flags timeout 600 python tools/scene_perf.py one_weekend:32 final_scene:32 stress:8 2>gpurun_out/r2e_perf.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'])"
done; done
