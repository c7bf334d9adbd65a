"""Closest-hit parity (north_star check 1 and 2): CUDA extend path vs the f64 oracle on the fixed synthetic ray set.

Primitive ids must match bit-exactly on every ray the oracle does not tag as ambiguous (a tie within
1e-9 relative, where the reference itself depends on its random BVH order);
hit distance, normal and (u, v) must agree within 1e-5.
"""
import numpy as np
import pytest

from common import check_trace_parity, make_ray_set

pytestmark = pytest.mark.gpu

CASES = [
    # scene, rays (N = 2^20 per scene), sampling box half-size around look_at (None = scene bounds),
    # aspect handed to Camera::new (None = the scene's own)
    ("cornel_box", 1 << 20, None, None),
    ("cornel_box", 1 << 18, None, 3840.0 / 2160.0),  # BASELINE config 4: the 16:9 camera sees past the room
    ("cornel_smoke", 1 << 20, None, None),
    ("simple_light", 1 << 20, 30.0, None),
    ("two_spheres", 1 << 20, 30.0, None),
    ("random_scene", 1 << 20, 15.0, None),
    ("one_weekend", 1 << 20, 15.0, None),  # BASELINE config 2 in its One-Weekend flavour (static spheres)
    ("final_scene", 1 << 20, 700.0, None),
    ("earth", 1 << 20, 10.0, None),
]


def _cornell_tie_reasons(rays, ot, amb):
    """Why the oracle calls a Cornell-box ray ambiguous: the bottom of the box lies IN the floor (main.rs:425-435 puts it at
    y = 0, main.rs:478-485 the floor too: a ray from above meets both at the same t and the reference's winner depends on
    its random BVH order), an edge of the room, or a silhouette (a 1e-9-relative nudge of the ray changes the primitive)."""
    a = amb != 0
    p = rays["origin"][a].astype(np.float64) + np.where(np.isfinite(ot[a]), ot[a], 0.0)[:, None] * rays["direction"][a].astype(np.float64)
    c, s = np.cos(np.radians(15.0)), np.sin(np.radians(15.0))
    q = p - np.array([265.0, 0.0, 295.0])
    loc = np.stack([c * q[:, 0] - s * q[:, 2], q[:, 1], s * q[:, 0] + c * q[:, 2]], axis=1)  # hittable.rs:241-245
    eps = 1e-5
    hit = np.isfinite(ot[a])
    floor = hit & (np.abs(p[:, 1]) < eps * 555) & (loc[:, 0] > -eps * 165) & (loc[:, 0] < 165 * (1 + eps)) & (loc[:, 2] > -eps * 165) & (loc[:, 2] < 165 * (1 + eps))
    room_edge = hit & ~floor & (((np.abs(p) < eps * 555) | (np.abs(p - 555) < eps * 555)).sum(axis=1) >= 2)
    return int(floor.sum()), int(room_edge.sum()), int(a.sum() - floor.sum() - room_edge.sum())


@pytest.mark.parametrize("name,n,extent,aspect", CASES)
def test_closest_hit_matches_oracle(rt, oracle, gpu_ctx, name, n, extent, aspect):
    api = rt.api
    hs = api.HostScene(name, seed=1)
    osc = oracle.OracleScene(hs.desc)
    gsc = api.Scene(gpu_ctx, hs.desc)
    prims = gsc.prims()
    assert len(prims) == osc.num_prims
    rays = make_ray_set(api, hs, osc, prims, n, extent, aspect=aspect)
    cache = {}
    check_trace_parity(gsc, osc, rays, label=f"{name}{' 16:9' if aspect else ''}", cache=cache)
    if name == "cornel_box":  # the ties have a name: say which, and that nothing else is excluded from the id comparison
        _, ot, _, _, _, amb = cache["trace"]
        floor, edge, silhouette = _cornell_tie_reasons(rays, ot, amb)
        print(f"[trace parity] cornel_box ties: {floor} box bottom in the floor, {edge} room edges, {silhouette} silhouettes / grazing of {n} rays")
        assert floor + edge + silhouette == int((amb != 0).sum()) and silhouette <= 2e-4 * n and edge <= 5e-4 * n
    gsc.close()


def _cornell_edge_points(rng, n):
    """Points on the 12 edges of the Cornell room and of the rotated tall box (main.rs:425-435,451-494)."""
    def box_edges(lo, hi, m):
        corner = lo + (hi - lo) * rng.integers(0, 2, (m, 3))
        axis = rng.integers(0, 3, m)
        pts = corner.astype(np.float64)
        pts[np.arange(m), axis] = lo[axis] + (hi[axis] - lo[axis]) * rng.random(m)
        return pts
    room = box_edges(np.zeros(3), np.full(3, 555.0), n // 2)
    local = box_edges(np.zeros(3), np.array([165.0, 330.0, 165.0]), n - n // 2)
    c, s = np.cos(np.radians(15.0)), np.sin(np.radians(15.0))  # hittable.rs:259-262 (object -> world), then Translate
    box = np.stack([c * local[:, 0] + s * local[:, 2], local[:, 1], -s * local[:, 0] + c * local[:, 2]], axis=1) + np.array([265.0, 0.0, 295.0])
    return np.concatenate([room, box])


@pytest.mark.parametrize("name", ["cornel_box", "cornel_smoke"])
def test_closest_hit_near_box_edges(rt, oracle, gpu_ctx, name):
    """Rays aimed within 1e-7 .. 1e-1 of the edges of the room and of the box: where the flat scan's face groups
    (one f32 slab computation names the entry and exit face of a box, kernels.cuh: closest_hit_flat) must fall back to
    per-face candidates.  Same bars as the synthetic ray set."""
    api = rt.api
    n = 1 << 17
    hs = api.HostScene(name, seed=1)
    osc = oracle.OracleScene(hs.desc)
    gsc = api.Scene(gpu_ctx, hs.desc)
    rng = np.random.Generator(np.random.Philox(0xED6E))
    target = _cornell_edge_points(rng, n)
    v = rng.normal(size=(n, 3))
    target += v / np.linalg.norm(v, axis=1, keepdims=True) * 10.0 ** rng.uniform(-7.0, -1.0, (n, 1))
    origin = np.where(rng.random((n, 1)) < 0.75, 5.0 + 545.0 * rng.random((n, 3)), np.array([[278.0, 278.0, -800.0]]))
    origin[n // 2:n // 2 + n // 8] = target[n // 2:n // 2 + n // 8] + np.array([[0.0, 0.0, -300.0]])  # axis-parallel rays
    rays = np.zeros(n, dtype=api.RAY_DTYPE)
    rays["origin"] = origin
    rays["direction"] = (target - origin) * 10.0 ** rng.uniform(-2.5, 0.5, (n, 1))
    rays["time"] = rng.random(n).astype(np.float32)
    gp, gt, gn, gff, guv = gsc.trace_closest(rays, seed=7)
    op, ot, on, off, ouv, amb = osc.trace_closest(rays, seed=7)
    keep = amb == 0
    assert keep.mean() > 0.9
    bad = keep & (gp != op)
    assert not bad.any(), f"{bad.sum()} primitive-id mismatches, first at ray {np.flatnonzero(bad)[:5]}: gpu {gp[bad][:5]} oracle {op[bad][:5]}"
    hit = keep & (op >= 0)
    assert hit.sum() > n // 2
    rel_t = np.abs(gt[hit].astype(np.float64) - ot[hit]) / np.maximum(np.abs(ot[hit]), 1e-30)
    assert rel_t.max() <= 1e-5, f"hit distance off by {rel_t.max():.3e} relative"
    assert np.abs(gn[hit].astype(np.float64) - on[hit]).max() <= 1e-5
    assert (gff[hit] == off[hit]).all()
    gsc.close()


@pytest.mark.parametrize("name", ["cornel_box", "cornel_smoke"])
def test_face_groups_change_nothing(rt, gpu_ctx, monkeypatch, name):
    """The face groups of the flat scan are a faster way to the same closest hits: with `RT1W_FACE_GROUPS=0` (one box per
    rectangle, read when the scene is committed) the same rays hit the same primitives at the same distances, and a
    render traces the same number of rays to the same image (up to the order of its fp32 additions)."""
    api = rt.api
    hs = api.HostScene(name, seed=1)
    grouped = api.Scene(gpu_ctx, hs.desc)
    monkeypatch.setenv("RT1W_FACE_GROUPS", "0")
    plain = api.Scene(gpu_ctx, hs.desc)
    monkeypatch.delenv("RT1W_FACE_GROUPS")
    rng = np.random.Generator(np.random.Philox(0xFACE))
    n = 1 << 16
    rays = np.zeros(n, dtype=api.RAY_DTYPE)
    rays["origin"] = -50.0 + 655.0 * rng.random((n, 3))
    v = rng.normal(size=(n, 3))
    rays["direction"] = v / np.linalg.norm(v, axis=1, keepdims=True) * rng.uniform(0.5, 20.0, (n, 1))
    rays["time"] = rng.random(n).astype(np.float32)
    a, b = grouped.trace_closest(rays, seed=3), plain.trace_closest(rays, seed=3)
    same = a[0] == b[0]
    # ties may resolve either way: two faces at a box edge, or the bottom of the Cornell box lying in the floor (a ray from
    # below meets both at the same t); everywhere else the primitive, the distance and the normal are identical
    assert same.mean() > 0.99
    assert np.array_equal(a[1][same], b[1][same]) and np.array_equal(a[2][same], b[2][same])
    tie = ~same
    assert ((a[0][tie] >= 0) & (b[0][tie] >= 0)).all()
    assert np.allclose(a[1][tie], b[1][tie], rtol=1e-6, atol=0.0)
    cam, p = hs.camera(), hs.params(width=96, height=96, spp=8, seed=5)
    ia, _, sa = grouped.render(cam, p)
    ib, _, sb = plain.render(cam, p)
    if name == "cornel_box":  # (a medium draws its free-flight number per candidate: the candidate sets differ, the estimate does not)
        assert sa.rays == sb.rays
        ok = np.isfinite(ia) & np.isfinite(ib)
        assert np.allclose(ia[ok], ib[ok], rtol=1e-3, atol=1e-3)
    else:
        assert abs(float(sa.rays) / float(sb.rays) - 1.0) < 0.02
    grouped.close(), plain.close()


def test_trace_empty_and_tiny(rt, oracle, gpu_ctx):
    api = rt.api
    hs = api.HostScene("cornel_box", seed=1)
    gsc = api.Scene(gpu_ctx, hs.desc)
    out = gsc.trace_closest(np.zeros(0, dtype=api.RAY_DTYPE))
    assert out[0].shape == (0,)
    # the known-answer centre ray: hits the box front at t = 109.15166604984
    r = np.zeros(1, dtype=api.RAY_DTYPE)
    r["origin"][0] = (278, 278, -800)
    r["direction"][0] = (0, 0, 10)
    prim, t, n, ff, uv = gsc.trace_closest(r)
    assert abs(t[0] - 109.15166604984) < 1e-4
    assert np.allclose(n[0], (-0.2588190451, 0, -0.9659258263), atol=1e-6)
    gsc.close()


def test_moving_spheres_at_extrapolated_times(rt, oracle, gpu_ctx):
    """Quirk 1 (main.rs:86): after a diffuse bounce the scattered ray's `time` is the hit parameter t, often >> 1, and
    `MovingSphere::center(time)` (moving_sphere.rs:23-26) extrapolates the sphere OUT of the [time0, time1] box its BVH
    nodes were built with (moving_sphere.rs:72-84).  Whether such a sphere is still found then depends on which boxes the
    ray meets on the way down - in the reference on its randomly built tree (bvh.rs:84), here on the leaf's own box.
    So: every ray whose closest hit involves no moving sphere on either side is answered exactly as ever; the rest are
    few, and the converged images agree (test_gpu_render_parity.py: random_scene, final_scene)."""
    api = rt.api
    hs = api.HostScene("random_scene", seed=1)
    osc = oracle.OracleScene(hs.desc)
    gsc = api.Scene(gpu_ctx, hs.desc)
    prims = gsc.prims()
    n = 1 << 17
    rays = make_ray_set(api, hs, osc, prims, n, 15.0)
    rng = np.random.Generator(np.random.Philox(77))
    rays["time"] = rng.uniform(0.0, 30.0, n).astype(np.float32)
    gp, gt, gn, gff, guv = gsc.trace_closest(rays, seed=3)
    op, ot, on, off, ouv, amb = osc.trace_closest(rays, seed=3)
    moving = np.array([p.kind == api.NODE_MOVING_SPHERE for p in prims])
    involved = (np.where(gp >= 0, moving[np.maximum(gp, 0)], False)) | (np.where(op >= 0, moving[np.maximum(op, 0)], False))
    keep = (amb == 0) & ~involved
    assert (gp[keep] == op[keep]).all()
    hit = keep & (op >= 0)
    assert (np.abs(gt[hit] - ot[hit]) <= 1e-5 * np.abs(ot[hit])).all()
    differ = (amb == 0) & involved & (gp != op)
    print(f"[trace parity] random_scene, ray times in [0, 30): {involved.mean():.4f} of the rays meet a moving sphere on either side, "
          f"{differ.mean():.5f} resolve differently (the sphere has left its [time0, time1] box)")
    assert differ.mean() <= 0.02
    gsc.close()


def test_global_primitives_stay_out_of_the_tree(rt, oracle, gpu_ctx, monkeypatch):
    """final_scene's fog - a ConstantMedium in a sphere of radius 5000 around everything (main.rs:734-745) - is in no BVH: its
    box contains every other primitive's, so every ray tests it once, after the traversal.  Same closest hits as with the
    sphere in the tree (`RT1W_GLOBAL_PRIMS=0`) and as the oracle's."""
    api = rt.api
    hs = api.HostScene("final_scene", seed=1)
    osc = oracle.OracleScene(hs.desc)
    out_of_tree = api.Scene(gpu_ctx, hs.desc)
    monkeypatch.setenv("RT1W_GLOBAL_PRIMS", "0")
    in_tree = api.Scene(gpu_ctx, hs.desc)
    monkeypatch.delenv("RT1W_GLOBAL_PRIMS")
    assert out_of_tree.info().n_global_prims == 1 and in_tree.info().n_global_prims == 0
    assert out_of_tree.info().n_bvh_nodes == in_tree.info().n_bvh_nodes - 2
    rays = make_ray_set(api, hs, osc, in_tree.prims(), 1 << 16, 700.0)
    cache = {}
    check_trace_parity(in_tree, osc, rays, label="final_scene, fog sphere in the tree", cache=cache)
    check_trace_parity(out_of_tree, osc, rays, label="final_scene, fog sphere out of the tree", cache=cache)
    a, b = out_of_tree.trace_closest(rays, seed=0x5EED), in_tree.trace_closest(rays, seed=0x5EED)
    same = a[0] == b[0]
    assert same[cache["trace"][5] == 0].all()
    for k in (1, 2, 3, 4):
        assert np.array_equal(a[k][same], b[k][same])
    out_of_tree.close(), in_tree.close()


@pytest.mark.parametrize("name,extent", [("final_scene", 700.0), ("cornel_smoke", None), ("random_scene", 15.0)])
def test_rays_starting_on_surfaces(rt, oracle, gpu_ctx, name, extent):
    """Rays that START on a primitive - the hit points of camera rays - heading outward or inward, and rays that start
    just INSIDE it and leave through the side they sit under at t = 1e-5 .. 1e-1 (around t_min = 0.001), with
    un-normalised directions: what the scattered rays of a render are.  A BVH leaf that holds an AABox leaves the box
    without testing its six sides (aabox.rs:84-103) when the f32 slab exit lies before t_min (kernels.cuh:
    box_first_sides): here that shortcut meets exits on either side of t_min and rays that enter the box through the
    side they start on.  Same bars as the synthetic ray set; the oracle's tie rule (a 1e-9-relative nudge of the ray
    changes the winner) takes out the rays whose start decides by rounding which side of the surface they are on."""
    api = rt.api
    n = 1 << 18
    hs = api.HostScene(name, seed=1)
    osc = oracle.OracleScene(hs.desc)
    gsc = api.Scene(gpu_ctx, hs.desc)
    base = make_ray_set(api, hs, osc, gsc.prims(), n, extent)
    prim, tt, normal, _, _, _ = osc.trace_closest(base, seed=1)
    ok = prim >= 0
    assert ok.mean() > 0.3
    idx = np.flatnonzero(ok)
    rng = np.random.Generator(np.random.Philox(0x51DE))
    pick = idx[rng.integers(0, len(idx), n)]
    o = base["origin"][pick].astype(np.float64) + tt[pick][:, None] * base["direction"][pick].astype(np.float64)
    nrm = normal[pick] / np.linalg.norm(normal[pick], axis=1, keepdims=True)
    v = rng.normal(size=(n, 3))
    tangent = v - (v * nrm).sum(1, keepdims=True) * nrm
    tangent /= np.linalg.norm(tangent, axis=1, keepdims=True)
    along = 10.0 ** rng.uniform(-3.0, 0.0, (n, 1)) * np.where(rng.random((n, 1)) < 0.5, 1.0, -1.0)
    d = (tangent * np.sqrt(np.maximum(1.0 - along * along, 0.0)) + nrm * along) * 10.0 ** rng.uniform(-1.0, 2.5, (n, 1))
    third = n // 3  # the last third starts under the surface: it reaches it at t_exit along d
    t_exit = 10.0 ** rng.uniform(-5.0, -1.0, (n, 1))
    o[2 * third:] -= (t_exit * d)[2 * third:]
    rays = np.zeros(n, dtype=api.RAY_DTYPE)
    rays["origin"] = o
    rays["direction"] = d
    rays["time"] = base["time"][pick]  # a moving sphere is where the camera ray met it
    check_trace_parity(gsc, osc, rays, label=f"{name}, rays starting on surfaces", max_ambiguous=0.3, min_hit_fraction=0.05)
    gsc.close()
