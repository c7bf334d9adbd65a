#!/bin/bash
mkdir -p gpurun_out
for b in lbvh sah; do
  echo "== builder $b"
  RT1W_BVH_BUILDER=$b RT1W_FLAGS=40 timeout 600 python tools/scene_perf.py stress:8 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'wide', d['wide_nodes'], d['wide_depth'], d['wide_children'], 'build', d['build_ms'])"
  RT1W_BVH_BUILDER=$b RT1W_FLAGS=40 RT1W_LIB=$PWD/raytracing-1w_b200/_build/variant_trav.so timeout 600 python tools/scene_perf.py stress:2 2>&1 >/dev/null | grep bvh | tail -1
  RT1W_BVH_BUILDER=$b RT1W_FLAGS=24 RT1W_LIB=$PWD/raytracing-1w_b200/_build/variant_trav.so timeout 600 python tools/scene_perf.py stress:2 2>&1 >/dev/null | grep bvh | tail -1
done
RT1W_NO_WARMUP=1 RT1W_FLAGS=40 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 1 -c 1 -f -o gpurun_out/r2g_stress_wide python tools/scene_perf.py stress:8 > gpurun_out/r2g_ncu_stress.log 2>&1
RT1W_NO_WARMUP=1 RT1W_FLAGS=20 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 1 -c 1 -f -o gpurun_out/r2g_final_binary python tools/scene_perf.py final_scene:64 > gpurun_out/r2g_ncu_final.log 2>&1
RT1W_NO_WARMUP=1 RT1W_FLAGS=20 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 1 -c 1 -f -o gpurun_out/r2g_oneweekend_binary python tools/scene_perf.py one_weekend:64 > gpurun_out/r2g_ncu_ow.log 2>&1
ls -la gpurun_out | grep r2g
