#!/bin/bash
# round 2, second GPU run: all GPU tests (no -x), durations, + ncu summaries of the BVH scenes before the wide BVH
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rP --durations=15 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2b_pytest.log | grep -v "^make\|^---"
for sc in final_scene:16 random_scene:32 stress:4; do
  name=${sc%%:*}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_wave" -s 4 -c 1 -f -o gpurun_out/r2b_${name} python tools/scene_perf.py $sc > gpurun_out/r2b_ncu_${name}.log 2>&1
done
ls -la gpurun_out | grep r2b
