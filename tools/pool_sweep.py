"""Wave capacity sweep: render time of a scene against rt1w_render_params.pool_paths.
    python tools/pool_sweep.py cornel_box 100 22 23 24 25"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
api = importlib.import_module("raytracing-1w_b200").api
name, spp = sys.argv[1], int(sys.argv[2])
ctx = api.Context(0)
hs = api.HostScene(name, seed=1)
sc = api.Scene(ctx, hs.desc)
cam = hs.camera()
for lg in map(int, sys.argv[3:]):
    best = 1e9
    for rep in range(3):
        _, _, st = sc.render(cam, hs.params(spp=spp, pool_paths=1 << lg))
        best = min(best, st.render_ms)
    print(f"{name} pool 2^{lg}: {best:.3f} ms, {st.paths / best / 1e3:.1f} Mpaths/s, {st.waves} waves, {st.launches} launches", flush=True)
