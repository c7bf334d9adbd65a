"""Pins the oracle's geometry / pdf / material functions to the known-answer table of SURVEY.md section 4
(values derived from the reference formulas; the reference itself ships no tests)."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle_binding import d3


def test_cornell_camera(rt, oracle):
    cam = rt.api.camera_new((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0)  # main.rs:888-892,940-951
    assert np.allclose(list(cam.w), (0, 0, -1)) and np.allclose(list(cam.u), (-1, 0, 0)) and np.allclose(list(cam.v), (0, 1, 0))
    assert np.allclose(list(cam.horizontal), (-7.2794046853, 0, 0), atol=1e-9)
    assert np.allclose(list(cam.vertical), (0, 7.2794046853, 0), atol=1e-9)
    assert np.allclose(list(cam.lower_left_corner), (281.6397023427, 274.3602976573, -790), atol=1e-9)
    o, d = d3(0, 0, 0), d3(0, 0, 0)
    oracle.load().oracle_camera_ray(C.byref(cam), 0.5, 0.5, o, d)
    assert np.allclose(list(o), (278, 278, -800)) and np.allclose(list(d), (0, 0, 10), atol=1e-9)  # un-normalised


def test_centre_ray_hits_rotated_box(rt, oracle):
    hs = rt.api.HostScene("cornel_box", seed=1)
    osc = oracle.OracleScene(hs.desc)
    prim, t, p, n, ff = osc.hit_one((278, 278, -800), (0, 0, 10))
    assert t == pytest.approx(109.15166604984, abs=1e-9)
    assert np.allclose(p, (278, 278, 291.5166604984), atol=1e-8)
    assert np.allclose(n, (-0.2588190451, 0, -0.9659258263), atol=1e-9)
    prims = rt.api.lower_prims(hs.desc)
    assert prims[prim].kind == rt.api.NODE_XY_RECT and prims[prim].p[4] == 0.0  # the box front, XYRect k = 0 in box space


def test_rotate_y_bbox(oracle):
    lo, hi = d3(0, 0, 0), d3(0, 0, 0)
    oracle.load().oracle_rotate_y_bbox(d3(0, 0, 0), d3(165, 330, 165), 15.0, lo, hi)
    assert np.allclose(list(lo), (0, 0, -42.7051424419), atol=1e-9)
    assert np.allclose(list(hi), (202.0829037796, 330, 159.3777613377), atol=1e-9)


def test_sphere_hit_and_pdfs(oracle):
    lib = oracle.load()
    o, c = np.array([278.0, 278, -800]), np.array([190.0, 90, 190])
    d = c - o
    t = lib.oracle_sphere_hit_t(d3(*c), 90.0, d3(*o), d3(*d), 0.001, math.inf)
    assert t == pytest.approx(1 - 90 / np.linalg.norm(d), abs=1e-12)
    assert t == pytest.approx(0.9110256568748, abs=1e-12)
    light = d3(213, 343, 227, 332, 554)
    assert lib.oracle_xz_rect_pdf_value(light, d3(278, 0, 279.5), d3(0, 1, 0)) == pytest.approx(554 ** 2 / 13650, rel=1e-12)
    assert lib.oracle_xz_rect_pdf_value(light, d3(100, 100, 100), d3(178, 454, 179.5)) == pytest.approx(22.6415418538, rel=1e-10)
    assert lib.oracle_xz_rect_pdf_value(light, d3(100, 100, 100), d3(0, -1, 0)) == 0.0
    assert lib.oracle_sphere_pdf_value(d3(190, 90, 190), 90.0, d3(190, 400, 190), d3(0, -1, 0)) == pytest.approx(3.6951624282, rel=1e-10)
    assert lib.oracle_sphere_pdf_value(d3(190, 90, 190), 90.0, d3(190, 400, 190), d3(0, 1, 0)) == 0.0


def test_dielectric_helpers(oracle):
    lib = oracle.load()
    assert lib.oracle_reflectance(1.0, 1 / 1.5) == pytest.approx(0.04)
    assert lib.oracle_reflectance(0.5, 1.5) == pytest.approx(0.04 + 0.96 * 0.5 ** 5)
    out = d3(0, 0, 0)
    s = 1 / math.sqrt(2)
    lib.oracle_refract(d3(s, -s, 0), d3(0, 1, 0), 1 / 1.5, out)
    assert np.allclose(list(out), (0.4714045208, -0.8819171037, 0), atol=1e-10)


def test_sphere_uv_and_onb(oracle):
    lib = oracle.load()
    uv = (C.c_double * 2)()
    for p, expect in [((1, 0, 0), (0.5, 0.5)), ((0, 1, 0), (0.5, 1.0)), ((0, 0, 1), (0.25, 0.5)), ((0, -1, 0), (0.5, 0.0)),
                      ((0, 0, -1), (0.75, 0.5))]:
        lib.oracle_sphere_uv(d3(*p), uv)
        assert np.allclose(list(uv), expect, atol=1e-12)
    o = (C.c_double * 9)()
    lib.oracle_onb_from_w(d3(0, 1, 0), o)
    assert np.allclose(list(o), (-1, 0, 0, 0, 0, -1, 0, 1, 0), atol=1e-12)
    lib.oracle_onb_from_w(d3(1, 0, 0), o)
    assert np.allclose(list(o), (0, -1, 0, 0, 0, 1, 1, 0, 0), atol=1e-12)
    lib.oracle_onb_from_w(d3(1, 1, 1), o)
    assert np.allclose(list(o)[:6], (-0.8164965809, 0.4082482905, 0.4082482905, 0, 0.7071067812, -0.7071067812), atol=1e-9)


def test_quantisation(rt, oracle):
    q = (C.c_int32 * 3)()
    oracle.load().oracle_quantise(d3(0.0, 0.25, 0.5), 1, q)
    assert list(q) == [0, 128, 181]
    oracle.load().oracle_quantise(d3(400.0, float("nan"), 100.0), 100, q)  # >= 1 -> 255; a NaN SUM -> 0 (color.rs:16-18)
    assert list(q) == [255, 0, 255]
    # the product's host-side resolve follows the same rule
    img = np.array([[[0.0, 0.25, 0.5], [400.0, np.nan, 100.0]]], dtype=np.float32)
    out1 = rt.api.resolve_rgb8(img[:, :1], 1)
    out2 = rt.api.resolve_rgb8(img[:, 1:], 100)
    assert out1.tolist() == [[[0, 128, 181]]] and out2.tolist() == [[[255, 0, 255]]]


def test_bvh_node_counts(oracle):
    """BVHNode::new shapes (bvh.rs:60-101): n objects -> nodes, depth."""
    lib = oracle.load()
    n, d = C.c_int32(), C.c_int32()
    for objects, nodes, depth in [(6, 7, 3), (8, 7, 3), (11, 13, 4), (400, 511, 9), (1000, 1023, 10)]:
        lib.oracle_bvh_count(objects, 3, C.byref(n), C.byref(d))
        assert (n.value, d.value) == (nodes, depth)


def test_perlin_and_textures(rt, oracle):
    hs = rt.api.HostScene("two_perlin_spheres", seed=1)
    osc = oracle.OracleScene(hs.desc)
    desc = hs.desc.contents
    tab = desc.perlins[0]
    lib = oracle.load()
    # noise is zero on the integer lattice (every weight vector is a lattice offset with a zero component sum weight)
    assert lib.oracle_perlin_noise(C.byref(tab), d3(3.0, -2.0, 7.0)) == pytest.approx(0.0, abs=1e-15)
    pts = np.random.default_rng(1).uniform(-50, 50, size=(200, 3))
    vals = np.array([lib.oracle_perlin_noise(C.byref(tab), d3(*p)) for p in pts])
    assert np.all(np.abs(vals) <= 1.0) and vals.std() > 0.1
    t7 = lib.oracle_perlin_turb(C.byref(tab), d3(1.3, 2.7, -0.4), 7)
    acc, w, p = 0.0, 1.0, np.array([1.3, 2.7, -0.4])
    for _ in range(7):
        acc += w * lib.oracle_perlin_noise(C.byref(tab), d3(*p))
        w *= 0.5
        p = p * 2
    assert t7 == pytest.approx(abs(acc), rel=1e-12)
    # NoiseTexture: 0.5 * (1 + sin(scale * z + 10 * turb(p, 7))) (texture.rs:57-65)
    noise_tex = [i for i in range(desc.n_textures) if desc.textures[i].type == rt.api.TEX_NOISE][0]
    c = osc.texture_value(noise_tex, 0.0, 0.0, (1.3, 2.7, -0.4))
    scale = desc.textures[noise_tex].scale
    assert c[0] == pytest.approx(0.5 * (1 + math.sin(scale * -0.4 + 10 * t7)), rel=1e-12) and c[0] == c[1] == c[2]


def test_checker_and_image_texture(rt, oracle):
    hs = rt.api.HostScene("two_spheres", seed=1)
    osc = oracle.OracleScene(hs.desc)
    desc = hs.desc.contents
    chk = [i for i in range(desc.n_textures) if desc.textures[i].type == rt.api.TEX_CHECKER][0]
    # sin(10x) sin(10y) sin(10z) < 0 -> odd (0.9, 0.9, 0.9), else even (0.2, 0.3, 0.1)  (texture.rs:46-55, main.rs:298-307)
    assert np.allclose(osc.texture_value(chk, 0, 0, (0.1, 0.1, 0.1)), (0.2, 0.3, 0.1))
    assert np.allclose(osc.texture_value(chk, 0, 0, (0.1, 0.1, -0.1)), (0.9, 0.9, 0.9))
    he = rt.api.HostScene("earth", seed=1)
    oe = oracle.OracleScene(he.desc)
    de = he.desc.contents
    img = [i for i in range(de.n_textures) if de.textures[i].type == rt.api.TEX_IMAGE][0]
    earth = rt.api.load_earthmap()
    h, w, _ = earth.shape
    for u, v in [(0.0, 1.0), (0.999999, 0.0), (0.5, 0.5), (1.5, -0.5), (0.25, 0.75)]:
        uu, vv = min(max(u, 0.0), 1.0), 1.0 - min(max(v, 0.0), 1.0)
        i, j = min(int(uu * w), w - 1), min(int(vv * h), h - 1)  # texture.rs:69-79
        assert np.allclose(oe.texture_value(img, u, v, (0, 0, 0)), earth[j, i] / 255.0)
