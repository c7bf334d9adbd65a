"""CPU-side numbers behind DESIGN.md section 10: what a ray through a scene costs in each acceleration structure.

    python tools/bvh_stats.py [scene[:spheres]] [rays]          # default: stress:1000000 300

Builds the host SAH tree and its 8-wide collapse over the scene's primitive boxes (rt1w_build_bvh_host - the structures scene
commit uploads, no GPU needed) and walks both in numpy for camera rays and for uniform rays inside the scene box (no closest-hit
shortening: the number of boxes a ray's slab test reaches, which is what a ray that hits nothing pays); beside them the cells and
candidate primitives of a uniform grid walked by a 3D-DDA at a few resolutions.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def decode(node):
    step = np.ldexp(1.0, node["exp"].astype(np.int64) - 127)
    o = node["origin"].astype(np.float64)[:, None]
    return o + node["qlo"].astype(np.float64) * step[:, None], o + node["qhi"].astype(np.float64) * step[:, None]


def slab_many(lo, hi, o, inv):  # lo, hi: [axis][k]
    a, b = (lo - o[:, None]) * inv[:, None], (hi - o[:, None]) * inv[:, None]
    return np.maximum(np.minimum(a, b).max(0), 0.0) <= np.maximum(a, b).min(0)


def walk_binary(nodes, o, inv):
    steps = leaves = 0
    stack = [0]
    while stack:
        nd = nodes[stack.pop()]
        left = int(nd["left_first"])
        pair = nodes[left:left + 2]
        steps += 1  # one 64-byte pair of children per step
        hit = slab_many(pair["min"].astype(np.float64).T, pair["max"].astype(np.float64).T, o, inv)
        for k in range(2):
            if hit[k]:
                if pair[k]["count"] == 0:
                    stack.append(left + k)
                else:
                    leaves += 1
    return steps, leaves


def walk_wide(wide, o, inv):
    visits = leaves = slots_hit = 0
    stack = [0]
    while stack:
        nd = wide[stack.pop()]
        visits += 1
        lo, hi = decode(nd)
        hit = slab_many(lo, hi, o, inv)
        imask, lmask = int(nd["imask"]), int(nd["leaf_mask"])
        ci = 0
        for s in range(8):
            if (imask >> s) & 1:
                if hit[s]:
                    stack.append(int(nd["child_base"]) + ci)
                    slots_hit += 1
                ci += 1
            elif (lmask >> s) & 1 and hit[s]:
                leaves += 1
                slots_hit += 1
    return visits, leaves, slots_hit


def grid_cost(lo, hi, rays, res):
    """Cells a 3D-DDA crosses and the primitive references it meets (a primitive is listed in every cell its box overlaps)."""
    gmin, gmax = lo.min(0), hi.max(0)
    cell = (gmax - gmin) / res
    c0 = np.clip(np.floor((lo - gmin) / cell).astype(np.int64), 0, res - 1)
    c1 = np.clip(np.floor((hi - gmin) / cell).astype(np.int64), 0, res - 1)
    count = np.zeros((res, res, res), dtype=np.int32)
    span = (c1 - c0 + 1)
    small = (span == 1).all(1)
    np.add.at(count, (c0[small, 0], c0[small, 1], c0[small, 2]), 1)
    for i in np.flatnonzero(~small):
        if span[i].prod() > 4096:  # room-sized primitives would stay in a small tree of their own
            continue
        count[c0[i, 0]:c1[i, 0] + 1, c0[i, 1]:c1[i, 1] + 1, c0[i, 2]:c1[i, 2] + 1] += 1
    cells = refs = 0
    for o, d in rays:
        inv = 1.0 / d
        a, b = (gmin - o) * inv, (gmax - o) * inv
        t0, t1 = max(np.minimum(a, b).max(), 0.0), np.maximum(a, b).min()
        if t0 > t1:
            continue
        p = o + (t0 + 1e-9) * d
        ijk = np.clip(np.floor((p - gmin) / cell).astype(np.int64), 0, res - 1)
        step = np.where(d > 0, 1, -1)
        nxt = gmin + (ijk + (d > 0)) * cell
        tmax = (nxt - o) * inv
        dt = np.abs(cell * inv)
        while True:
            cells += 1
            refs += int(count[ijk[0], ijk[1], ijk[2]])
            k = int(np.argmin(tmax))
            ijk[k] += step[k]
            if ijk[k] < 0 or ijk[k] >= res or tmax[k] > t1:
                break
            tmax[k] += dt[k]
    return cells / len(rays), refs / len(rays), float(count.mean())


def main():
    api = importlib.import_module("raytracing-1w_b200").api
    spec = sys.argv[1] if len(sys.argv) > 1 else "stress:1000000"
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    name, _, arg = spec.partition(":")
    hs = api.HostScene(name, seed=1, **({"stress_spheres": int(arg or 1000000)} if name == "stress" else {}))
    prims = api.lower_prims(hs.desc)
    lo = np.array([list(p.bbox_min) for p in prims])
    hi = np.array([list(p.bbox_max) for p in prims])
    ok = np.isfinite(lo).all(1) & np.isfinite(hi).all(1)
    lo, hi = lo[ok], hi[ok]
    t = api.build_bvh_host(lo, hi)
    print(f"{name}: {len(lo)} primitive boxes, binary tree {len(t['nodes'])} nodes (depth {t['depth']}), 8-wide tree {len(t['wide_nodes'])} nodes "
          f"(depth {t['wide_depth']})")
    rng = np.random.Generator(np.random.Philox(3))
    cam = hs.camera()
    origin = np.array(list(cam.origin))
    llc, hor, ver = (np.array(list(v)) for v in (cam.lower_left_corner, cam.horizontal, cam.vertical))
    st = rng.random((n_rays, 2))
    sets = {"camera rays": [(origin, llc + s * hor + u * ver - origin) for s, u in st]}
    gmin, gmax = lo.min(0), hi.max(0)
    v = rng.normal(size=(n_rays, 3))
    sets["uniform rays in the scene box"] = [(gmin + rng.random(3) * (gmax - gmin), d / np.linalg.norm(d)) for d in v]
    for label, rays in sets.items():
        rays = [(o, np.where(np.abs(d) < 1e-9, 1e-9, d)) for o, d in rays]
        b = np.array([walk_binary(t["nodes"], o, 1.0 / d) for o, d in rays]).mean(0)
        w = np.array([walk_wide(t["wide_nodes"], o, 1.0 / d) for o, d in rays]).mean(0)
        print(f"  {label}: binary {b[0]:.1f} node steps, {b[1]:.2f} leaf boxes reached | 8-wide {w[0]:.1f} node visits "
              f"({w[2] / max(w[0], 1e-9):.2f} slots hit per visit), {w[1]:.2f} leaf slots reached")
        for res in (64, 100, 128):
            c, r, occ = grid_cost(lo, hi, rays, res)
            print(f"      grid {res}^3 ({occ:.2f} references per cell): {c:.0f} cells, {r:.1f} primitive references per ray")


if __name__ == "__main__":
    main()
