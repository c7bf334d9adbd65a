#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_gpu_stress_parity.py > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2m_pytest.log | grep -v "^make\|^---"
for flags in 0 64; do
  echo "== flags $flags (64 = one launch per wave)"
  RT1W_FLAGS=$flags timeout 600 python tools/scene_perf.py cornel_box:100 cornel_box:12 cornel_smoke:64 one_weekend:32 random_scene:32 final_scene:32 2>gpurun_out/r2m_perf.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'waves', d['waves'])"
done
python bench.py --no-configs --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench C1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'roofline', d['roofline']['frac'])"
