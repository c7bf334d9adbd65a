#!/bin/bash
# A/B: the fused wave kernel (the product) against the same wave as a split pipeline - a generate + shade launch that leaves its
# rays in HBM and an extend launch that picks them up (RT1W_SPLIT_PIPELINE=1; render.cu: k_wave<.., PHASE>) - north_star (b)'s
# shape at the same tuning level.  Prints throughput both ways and checks that the two produce the same image.
OUT=${1:-gpurun_out/ab_split.txt}
mkdir -p gpurun_out
{
for mode in fused split; do
  echo "== $mode"
  for i in 1 2; do
    env $( [ $mode = split ] && echo RT1W_SPLIT_PIPELINE=1 || echo RT1W_NOOP=1 ) timeout 600 python tools/scene_perf.py cornel_box:100 one_weekend:32 final_scene:32 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['scene'], d['image'], d['spp'], 'spp:', d['render_ms'], 'ms', d['mpaths_s'], 'Mpaths/s', d['mrays_s'], 'Mrays/s', d['waves'], 'waves')"
  done
  env $( [ $mode = split ] && echo RT1W_SPLIT_PIPELINE=1 || echo RT1W_NOOP=1 ) python - <<PY
import importlib, numpy as np
api = importlib.import_module("raytracing-1w_b200").api
ctx = api.Context(0)
for name in ("cornel_box", "one_weekend", "final_scene"):
    hs = api.HostScene(name, seed=1); sc = api.Scene(ctx, hs.desc)
    img, _, st = sc.render(hs.camera(), hs.params(width=128, spp=16, seed=3))
    np.save(f"gpurun_out/ab_split_{name}_$mode.npy", img); print(" ", name, "128 px x 16 spp:", st.rays, "rays,", st.launches, "launches")
PY
done
python - <<PY
import numpy as np
for name in ("cornel_box", "one_weekend", "final_scene"):
    a, b = np.load(f"gpurun_out/ab_split_{name}_fused.npy"), np.load(f"gpurun_out/ab_split_{name}_split.npy")
    ok = np.isfinite(a) & np.isfinite(b)
    print(name, "same image:", bool((np.isfinite(a) == np.isfinite(b)).all() and np.allclose(a[ok], b[ok], rtol=1e-3, atol=1e-3)))
PY
rm -f gpurun_out/ab_split_*.npy
} 2>&1 | tee $OUT
