#!/bin/bash
mkdir -p gpurun_out
export RT1W_NO_WARMUP=1
RT1W_FLAGS=24 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 2 -c 1 -f -o gpurun_out/r2o_final_persistent python tools/scene_perf.py final_scene:32 > gpurun_out/r2o_ncu.log 2>&1
RT1W_FLAGS=24 RT1W_LIB=$PWD/raytracing-1w_b200/_build/variant_trav.so timeout 600 python tools/scene_perf.py final_scene:8 2>&1 >/dev/null | grep bvh | tail -1
