import importlib, sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
rt = importlib.import_module("raytracing-1w_b200"); api = rt.api
import oracle_binding as ob
from common import make_ray_set
name = sys.argv[1] if len(sys.argv) > 1 else "cornel_box"
hs = api.HostScene(name, seed=1)
ctx = api.Context(0); scene = api.Scene(ctx, hs.desc); osc = ob.OracleScene(hs.desc)
prims = scene.prims()
rays = make_ray_set(api, hs, osc, prims, 1 << 14)
gp, gt, gn, gff, guv = scene.trace_closest(rays, seed=1)
op, ot, on, off, ouv, amb = osc.trace_closest(rays, seed=1)
keep = amb == 0
bad = np.flatnonzero(keep & (gp != op))
print("n", len(rays), "amb", amb.mean(), "bad", len(bad))
for i in bad[:12]:
    print(i, rays[i], "gpu", gp[i], gt[i], "oracle", op[i], ot[i], "kinds", prims[gp[i]].kind if gp[i] >= 0 else None, prims[op[i]].kind if op[i] >= 0 else None)
info = scene.info(); print(info.n_prims, info.n_bvh_nodes, info.bvh_depth)
