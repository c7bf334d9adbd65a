#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_libs.sh "cornel_box:100 final_scene:32 random_scene:32:1200 one_weekend:32 stress:8" librt1w variant_fsc
RT1W_LIB=$PWD/raytracing-1w_b200/_build/variant_fsc.so python -m pytest tests/test_gpu_shading_hooks.py -m gpu -x -q > gpurun_out/s2j_pytest.log 2>&1; echo "pytest(fsc) exit $?"; tail -2 gpurun_out/s2j_pytest.log
