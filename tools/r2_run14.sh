#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_shading_hooks.py tests/test_gpu_adhoc_scenes.py -m gpu -q --tb=short -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2n_pytest.log | grep -v "^make\|^---"
timeout 600 python tools/scene_perf.py cornel_box:100 one_weekend:32 random_scene:32:1200 final_scene:32 two_perlin_spheres:64:1200 earth:64:1200 2>gpurun_out/r2n_perf.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['image'], d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'waves', d['waves'])"
