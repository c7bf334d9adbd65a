#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_gpu_stress_parity.py > gpurun_out/r2p_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2p_pytest.log | grep -v "^make\|^---"
nvidia-smi --query-gpu=memory.used --format=csv,noheader
timeout 600 python tools/scene_perf.py cornel_box:100 cornel_smoke:64 one_weekend:32 random_scene:32:1200 final_scene:32 stress:8 2>gpurun_out/r2p_perf.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['image'], d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'waves', d['waves'])"
python - <<'PY'
import importlib, subprocess
api = importlib.import_module("raytracing-1w_b200").api
hs = api.HostScene("cornel_box", seed=1); ctx = api.Context(0); sc = api.Scene(ctx, hs.desc)
sc.render(hs.camera(), hs.params())
print("device memory in use with the Cornell pool:", subprocess.check_output(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader"], text=True).strip())
PY
