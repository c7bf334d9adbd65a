#!/bin/bash
# round 2, first GPU run: the GPU tests (new parity cases) and the throughput of every BASELINE config before the wide BVH
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
nproc
python -m pytest tests -m gpu -q --tb=short -rP -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2a_pytest.log
grep -h "trace parity\|stress " gpurun_out/r2a_pytest.log | head -40
timeout 600 python tools/scene_perf.py cornel_box:100 random_scene:64 one_weekend:64 final_scene:64 cornel_smoke:64 stress:16 > gpurun_out/r2a_scene_perf.json 2> gpurun_out/r2a_scene_perf.err; cat gpurun_out/r2a_scene_perf.json
RT1W_LIB=$PWD/raytracing-1w_b200/_build/variant_trav.so timeout 600 python tools/scene_perf.py random_scene:8 final_scene:8 stress:2 > gpurun_out/r2a_trav.json 2> gpurun_out/r2a_trav.err; grep bvh gpurun_out/r2a_trav.err
