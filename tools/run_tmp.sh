#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/s3a_pytest.log 2>&1; echo "pytest exit $? after $((SECONDS - T0)) s"; tail -2 gpurun_out/s3a_pytest.log
run() { echo "== $*"; env "$@" python tools/scene_perf.py $SC 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['image'], d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'waves', d['waves'])"; }
SC="stress:8 stress:32 cornel_box:100 cornel_box:12 final_scene:32 random_scene:32:1200"
run RT1W_FLAGS=0
run RT1W_FLAGS=64
run RT1W_TAIL_FACTOR=0.5
run RT1W_TAIL_FACTOR=2
run RT1W_TAIL_FACTOR=4
run RT1W_TAIL_FACTOR=8
run RT1W_FLAGS=0
echo "perf done after $((SECONDS - T0)) s"
python tools/pool_sweep.py cornel_box 100 23 24 25
python tools/pool_sweep.py final_scene 32 23 24
python tools/pool_sweep.py random_scene 32 23 24
echo "done after $((SECONDS - T0)) s"
