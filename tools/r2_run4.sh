#!/bin/bash
# round 2, fourth GPU run: wide BVH correctness + A/B of tree layout x wave kernel on the BVH scenes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_wide_bvh.py tests/test_gpu_multi.py tests/test_gpu_shading_hooks.py tests/test_gpu_stress_parity.py tests/test_gpu_lbvh.py -m gpu -q --tb=short -rP -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2d_pytest.log | grep -v "^make\|^---"; grep "wide bvh" gpurun_out/r2d_pytest.log
for flags in 20 24 36 40; do   # 16 binary / 32 wide  +  4 lockstep / 8 persistent
  echo "== flags $flags"
  RT1W_FLAGS=$flags timeout 600 python tools/scene_perf.py random_scene:32 one_weekend:32 final_scene:32 stress:8 2>gpurun_out/r2d_perf_$flags.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'wide', d['wide_nodes'], d['wide_depth'], d['wide_children'], 'build_ms', d['build_ms'])"
done
for flags in 24 40; do
  RT1W_FLAGS=$flags RT1W_LIB=$PWD/raytracing-1w_b200/_build/variant_trav.so timeout 600 python tools/scene_perf.py final_scene:4 stress:2 > /dev/null 2> gpurun_out/r2d_trav_$flags.err; echo "flags $flags"; grep bvh gpurun_out/r2d_trav_$flags.err | tail -2
done
