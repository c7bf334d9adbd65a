#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_wide_bvh.py -m gpu -q --tb=short -rP -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2f_pytest.log | grep -v "^make\|^---"; grep "wide bvh" gpurun_out/r2f_pytest.log
for lib in librt1w variant_isw6 variant_isw12 variant_ll12; do
for flags in 36 40; do
  echo "== $lib flags $flags"
  RT1W_LIB=$PWD/raytracing-1w_b200/_build/$lib.so RT1W_FLAGS=$flags timeout 600 python tools/scene_perf.py one_weekend:32 final_scene:32 stress:8 2>gpurun_out/r2f_perf.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'wide', d['wide_nodes'], d['wide_depth'], d['wide_children'], 'build', d['build_ms'])"
done; done
RT1W_FLAGS=40 RT1W_LIB=$PWD/raytracing-1w_b200/_build/variant_trav.so timeout 600 python tools/scene_perf.py one_weekend:4 final_scene:4 stress:2 > /dev/null 2> gpurun_out/r2f_trav.err; grep bvh gpurun_out/r2f_trav.err
