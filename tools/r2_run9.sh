#!/bin/bash
mkdir -p gpurun_out
for lib in librt1w variant_pb3 variant_pb5 variant_pb6; do
for flags in 40 24; do
  echo "== $lib flags $flags"
  RT1W_LIB=$PWD/raytracing-1w_b200/_build/$lib.so RT1W_FLAGS=$flags timeout 600 python tools/scene_perf.py stress:8 2>gpurun_out/r2i_perf.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'])"
done; done
RT1W_FLAGS=20 timeout 600 python tools/scene_perf.py cornel_box:100 one_weekend:32 random_scene:32 final_scene:32 2>gpurun_out/r2i_perf.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], 'ms', d['render_ms'], 'Mpaths/s', d['mpaths_s'], 'Mrays/s', d['mrays_s'])"
