#!/bin/bash
for i in 1 2; do
timeout 600 python tools/scene_perf.py cornel_box:100 cornel_smoke:64 one_weekend:32 final_scene:32 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['image'], d['spp'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'], 'waves', d['waves'])"
done
