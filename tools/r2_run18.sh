#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_wide_bvh.py -m gpu -q --tb=short -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2r_pytest.log | grep -v "^make\|^---"
for lib in librt1w variant_st16 variant_dbl; do
  echo "== $lib"
  for i in 1 2; do
  RT1W_LIB=$PWD/raytracing-1w_b200/_build/$lib.so timeout 600 python tools/scene_perf.py stress:8 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'])"
  done
done
