#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2a_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/s2a_pytest.log
fmt='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(" ", d["scene"], d["image"], d["spp"], "spp:", d["render_ms"], "ms", d["mpaths_s"], "Mpaths/s", d["mrays_s"], "Mrays/s", d["waves"], "waves")'
for i in 1 2; do
python tools/scene_perf.py cornel_box:100 random_scene:32 one_weekend:32 final_scene:32 cornel_smoke:32 stress:8 2>/dev/null | python -c "$fmt"
done
