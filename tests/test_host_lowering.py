"""Host-side logic above and below the C ABI that needs no GPU: the C++ mirror of the scene functions, the
lowering (wrapper chains, box expansion, primitive ids), its error behaviour, and the output writer."""
import ctypes as C
import math

import numpy as np
import pytest


def _counts(api, prims):
    kinds = {}
    for p in prims:
        kinds[p.kind] = kinds.get(p.kind, 0) + 1
    return kinds


def test_cornel_box_lowering(rt, oracle):
    api = rt.api
    hs = api.HostScene("cornel_box", seed=1)
    s = hs.settings
    assert (s.image_width, s.image_height, s.samples_per_pixel, s.max_depth) == (600, 600, 100, 50)  # main.rs:868-870,801
    assert list(s.background) == [0, 0, 0] and list(s.look_from) == [278, 278, -800]
    prims = api.lower_prims(hs.desc)
    assert len(prims) == 13  # 5 walls + light + 6 box sides + glass sphere
    kinds = _counts(api, prims)
    assert kinds == {api.NODE_YZ_RECT: 4, api.NODE_XZ_RECT: 5, api.NODE_XY_RECT: 3, api.NODE_SPHERE: 1}
    light = prims[2]
    assert light.kind == api.NODE_XZ_RECT and light.flags == 1 and list(light.p)[:5] == [213, 343, 227, 332, 554]  # FlipFace, main.rs:470-477
    box = prims[6:12]
    assert all(p.frame == 0 for p in box) and all(p.frame == -1 for p in prims[:6])
    # world bounds of the rotated + translated box
    lo = np.min([list(p.bbox_min) for p in box], axis=0)
    hi = np.max([list(p.bbox_max) for p in box], axis=0)
    assert np.allclose(lo, (265, 0, 252.2948575581), atol=2e-4) and np.allclose(hi, (467.0829037796, 330, 454.3777613377), atol=2e-4)
    desc = hs.desc.contents
    assert desc.has_lights == 1 and desc.n_lights == 2
    assert oracle.OracleScene(hs.desc).num_prims == 13


def test_face_groups(rt):
    """The flat scan's face groups (api.cu: find_face_groups): the five walls of the Cornell room and the six sides of
    the box (aabox.rs:29-76 order: XY@z1, XY@z0, XZ@y1, XZ@y0, YZ@x1, YZ@x0) each share one slab computation."""
    api = rt.api
    hs = api.HostScene("cornel_box", seed=1)  # (the description lives as long as the host scene)
    group, face, n = api.lower_face_groups(hs.desc)
    assert n == 2
    # prims: YZ@555, YZ@0, light, XZ@0, XZ@555, XY@555 (main.rs:451-494), box sides, sphere
    assert list(group) == [0, 0, -1, 0, 0, 0, 1, 1, 1, 1, 1, 1, -1]
    assert list(face) == [1, 0, -1, 2, 3, 5, 5, 4, 3, 2, 1, 0, -1]
    hs = api.HostScene("cornel_smoke", seed=1)
    group, face, n = api.lower_face_groups(hs.desc)
    assert n == 1 and list(group[:6]) == [0, 0, -1, 0, 0, 0] and (group[6:] == -1).all()  # the smoke boxes are media, not rectangles
    for name in ("two_spheres", "simple_light", "earth"):  # fewer than three faces of a common box: no group
        hs = api.HostScene(name, seed=1)
        group, face, n = api.lower_face_groups(hs.desc)
        assert n == 0 and (group == -1).all()


@pytest.mark.parametrize("name,n_prims", [("two_spheres", 2), ("two_perlin_spheres", 2), ("earth", 1), ("simple_light", 3),
                                          ("cornel_smoke", 8), ("final_scene", 3409)])
def test_scene_primitive_counts(rt, oracle, name, n_prims):
    hs = rt.api.HostScene(name, seed=1)
    prims = rt.api.lower_prims(hs.desc)
    assert len(prims) == n_prims
    assert oracle.OracleScene(hs.desc).num_prims == n_prims
    for p in prims:
        assert all(math.isfinite(v) for v in list(p.bbox_min) + list(p.bbox_max))


def test_random_scene_is_seeded(rt):
    api = rt.api
    a, b, c = api.HostScene("random_scene", seed=1), api.HostScene("random_scene", seed=1), api.HostScene("random_scene", seed=2)
    pa, pb, pc = (api.lower_prims(s.desc) for s in (a, b, c))
    assert 400 < len(pa) <= 488
    assert [list(p.p) for p in pa] == [list(p.p) for p in pb]
    assert [list(p.p) for p in pa] != [list(p.p) for p in pc]
    assert a.settings.samples_per_pixel == 500 and a.settings.aperture == 0.1  # main.rs:816-827
    kinds = _counts(api, pa)
    assert kinds[api.NODE_MOVING_SPHERE] > 250  # ~80 % of the small spheres are moving lambertians (main.rs:220-237)


def test_cornel_smoke_media(rt):
    api = rt.api
    hs = api.HostScene("cornel_smoke", seed=1)  # keep the owner of the description alive
    prims = api.lower_prims(hs.desc)
    media = [p for p in prims if p.kind == api.NODE_CONSTANT_MEDIUM]
    assert len(media) == 2 and all(p.boundary == api.NODE_AABOX and p.frame >= 0 for p in media)
    assert sorted(round(-1.0 / p.p[3], 6) for p in media) == [0.01, 0.01]  # density 0.01 (main.rs:563-577)


def test_lowering_errors_mirror_the_reference_panics(rt):
    api = rt.api
    b = api.DescBuilder()
    m = b.lambertian(b.solid(0.5, 0.5, 0.5))
    b.set_world(b.bvh([]))  # BVHNode::new panics on an empty list (bvh.rs:61)
    with pytest.raises(api.Rt1wError) as e:
        api.lower_prims(b.desc())
    assert e.value.status == api.ERR_INVALID and "empty" in str(e.value)

    b = api.DescBuilder()
    m = b.lambertian(b.solid(0.5, 0.5, 0.5))
    b.set_world(b.bvh([b.sphere((0, 0, 0), 1.0, 99)]))  # dangling material handle
    with pytest.raises(api.Rt1wError):
        api.lower_prims(b.desc())

    b = api.DescBuilder()
    m = b.lambertian(b.solid(0.5, 0.5, 0.5))
    b.set_world(b.bvh([b.sphere((0, 0, 0), 1.0, m)]))
    b.set_lights([])  # Some(vec![]) -> choose().unwrap() panics (hittable.rs:153)
    with pytest.raises(api.Rt1wError) as e:
        api.lower_prims(b.desc())
    assert e.value.status == api.ERR_INVALID

    b = api.DescBuilder()
    m = b.lambertian(b.solid(0.5, 0.5, 0.5))
    inner = b.bvh([b.sphere((0, 0, 0), 1.0, m), b.sphere((3, 0, 0), 1.0, m)])
    b.set_world(b.bvh([b.constant_medium(inner, 0.1, b.solid(1, 1, 1))]))  # a two-object boundary is not lowered
    with pytest.raises(api.Rt1wError) as e:
        api.lower_prims(b.desc())
    assert e.value.status == api.ERR_UNSUPPORTED


def test_wrapper_chain_frames(rt, oracle):
    """Nested wrappers collapse to one frame whose local ray equals the oracle's nested transforms."""
    api = rt.api
    b = api.DescBuilder()
    m = b.metal((0.8, 0.8, 0.8), 0.0)
    box = b.aabox((0, 0, 0), (10, 20, 30), m)
    node = b.translate(b.rotate_y(b.translate(b.rotate_y(box, 25.0), (5, 1, -2)), -40.0), (100, 0, 50))
    b.set_world(b.bvh([node, b.flip_face(b.xz_rect(-5, 5, -5, 5, 60, m))]))
    desc = b.desc()
    prims = api.lower_prims(desc)
    assert len(prims) == 7 and len({p.frame for p in prims[:6]}) == 1 and prims[6].frame == -1 and prims[6].flags == 1
    osc = oracle.OracleScene(desc)
    # a ray aimed at the centre of the transformed box must hit one of its sides; the lowered world-space bounds contain the hit
    lo = np.min([list(p.bbox_min) for p in prims[:6]], axis=0)
    hi = np.max([list(p.bbox_max) for p in prims[:6]], axis=0)
    centre = 0.5 * (lo + hi)
    o = centre + np.array([300.0, 40.0, 120.0])
    prim, t, p, n, ff = osc.hit_one(o, centre - o)
    assert 0 <= prim < 6
    assert np.all(np.array(p) >= lo - 1e-6) and np.all(np.array(p) <= hi + 1e-6)


def test_ppm_writer_and_resolve(rt, tmp_path):
    api = rt.api
    img = np.zeros((2, 3, 3), dtype=np.float32)
    img[0, 0] = (1.0, 0.25, 0.0)
    img[1, 2] = (4.0, 4.0, 4.0)
    rgb8 = api.resolve_rgb8(img, 4)  # means 0.25, 0.0625, 0 -> sqrt -> 128, 64, 0 ; 1.0 -> clamp 0.999 -> 255
    assert rgb8[0, 0].tolist() == [128, 64, 0] and rgb8[1, 2].tolist() == [255, 255, 255]
    path = str(tmp_path / "out.ppm")
    api.write_ppm(path, rgb8)
    lines = open(path).read().split("\n")
    assert lines[:3] == ["P3", "3 2", "255"]  # main.rs:953
    assert lines[3] == "128 64 0" and lines[8] == "255 255 255" and len([l for l in lines if l]) == 3 + 6


def test_camera_aspect_and_lens(rt):
    api = rt.api
    hs = api.HostScene("random_scene", seed=1)
    cam = hs.camera()
    assert cam.lens_radius == pytest.approx(0.05) and (cam.time0, cam.time1) == (0.0, 1.0)  # aperture 0.1 / 2 (camera.rs:57)
    cam2 = hs.camera(aspect=2.0)
    assert np.linalg.norm(list(cam2.horizontal)) == pytest.approx(2.0 * np.linalg.norm(list(cam2.vertical)))
