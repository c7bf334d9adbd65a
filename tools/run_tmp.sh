#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2b_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/s2b_pytest.log
bash tools/ab_libs.sh "final_scene:32 random_scene:32:1200 one_weekend:32 cornel_box:100" librt1w
echo "== counts: lockstep (default) then persistent (flags 8|16)"
for f in 0 24; do
RT1W_FLAGS=$f RT1W_LIB=$PWD/raytracing-1w_b200/_build/variant_trav.so python tools/scene_perf.py final_scene:8 random_scene:8:1200 one_weekend:8 2>&1 | grep -a "^\[bvh\]\|mrays" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'flags', d['flags'], 'Mrays/s', d['mrays_s'])
    else: print(l.strip())"
done
