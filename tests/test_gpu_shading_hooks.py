"""Pointwise parity of the DEVICE shading functions (rt1w_eval_* hooks): the functions the wave kernels call for light
pdfs, textures, Perlin noise, the dielectric helpers and `Material::scatter`, held value by value against SURVEY.md
section 4's known answers (derived from the reference formulas) and against the oracle at 10^5 random inputs.

Tolerances (the device computes these in f32 on f64-relative coordinates):
  light pdfs 2e-5 relative on the known answers, 1e-4 relative on random inputs (silhouette flips aside);
  refract / reflect 1e-6 absolute, reflectance 1e-6; Perlin noise 1e-5, turb(7) 2e-5, NoiseTexture 3e-4 absolute;
  checker and image lookups equal except within rounding of a cell border; scattered directions 1e-4 of their length;
  scatter weights 1e-3 relative.
"""
import ctypes as C
import math

import numpy as np
import pytest

from common import make_ray_set
from oracle_binding import d3

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cornell(rt, oracle, gpu_ctx):
    api = rt.api
    hs = api.HostScene("cornel_box", seed=1)
    gsc = api.Scene(gpu_ctx, hs.desc)
    osc = oracle.OracleScene(hs.desc)
    yield api, hs, gsc, osc
    gsc.close()


LIGHT = (213.0, 343.0, 227.0, 332.0, 554.0)   # main.rs:873-879
GLASS = ((190.0, 90.0, 190.0), 90.0)          # main.rs:880-886


def test_light_pdf_known_answers(cornell):
    """aarect.rs:119-138 and sphere.rs:72-90 on the device: SURVEY.md section 4's table."""
    api, hs, gsc, osc = cornell
    o = np.array([[278, 0, 279.5], [100, 100, 100], [100, 100, 100]], dtype=np.float64)
    v = np.array([[0, 1, 0], [178, 454, 179.5], [0, -1, 0]], dtype=np.float32)
    rect = gsc.eval_light_pdf(o, v, light=0)
    assert rect[0] == pytest.approx(22.4846886447, rel=2e-5) and rect[1] == pytest.approx(22.6415418538, rel=2e-5) and rect[2] == 0.0
    o2 = np.array([[190, 400, 190], [190, 400, 190]], dtype=np.float64)
    v2 = np.array([[0, -1, 0], [0, 1, 0]], dtype=np.float32)
    sph = gsc.eval_light_pdf(o2, v2, light=1)
    assert sph[0] == pytest.approx(3.6951624282, rel=2e-5) and sph[1] == 0.0
    # the list's own pdf_value: the average over its entries (hittable.rs:144-150)
    both = gsc.eval_light_pdf(np.concatenate([o, o2]), np.concatenate([v, v2]))
    assert both[0] == pytest.approx(0.5 * 22.4846886447, rel=2e-5) and both[3] == pytest.approx(0.5 * 3.6951624282, rel=2e-5)


def test_light_pdf_random_points(cornell, oracle):
    api, hs, gsc, osc = cornell
    lib = oracle.load()
    rng = np.random.Generator(np.random.Philox(21))
    n = 100_000
    o = rng.uniform(1.0, 554.0, (n, 3))
    target = np.where(rng.random((n, 1)) < 0.5, np.array([[278.0, 554.0, 279.5]]), np.array([GLASS[0]])) + rng.normal(size=(n, 3)) * 80.0
    v = ((target - o) * rng.uniform(0.01, 2.0, (n, 1))).astype(np.float32)
    got = [gsc.eval_light_pdf(o, v, light=k) for k in (0, 1)]
    rect5, c3 = d3(*LIGHT), d3(*GLASS[0])
    want = np.empty((2, n))
    for i in range(n):
        oi, vi = d3(*o[i]), d3(*v[i].astype(np.float64))
        want[0, i] = lib.oracle_xz_rect_pdf_value(rect5, oi, vi)
        want[1, i] = lib.oracle_sphere_pdf_value(c3, GLASS[1], oi, vi)
    for k in (0, 1):
        assert (want[k] > 0).mean() > 0.1
        flip = (got[k] > 0) != (want[k] > 0)  # the ray passes the light's silhouette within f32 rounding
        assert flip.mean() <= 1e-3, f"light {k}: {flip.sum()} hit / miss disagreements"
        ok = ~flip & (want[k] > 0) & np.isfinite(want[k])
        rel = np.abs(got[k][ok] - want[k][ok]) / want[k][ok]
        assert np.quantile(rel, 0.999) <= 1e-4 and rel.max() <= 1e-2, f"light {k}: pdf off by {rel.max():.2e} relative"
    mix = gsc.eval_light_pdf(o, v)
    both = 0.5 * (got[0].astype(np.float64) + got[1].astype(np.float64))
    fin = np.isfinite(both)
    assert (np.isfinite(mix) == fin).all() and np.abs(mix[fin] - both[fin]).max(initial=0.0) <= 1e-5 * both[fin].max(initial=1.0)


def test_dielectric_helpers_on_device(gpu_ctx, oracle):
    """material.rs:94-96,114-125 on the device: SURVEY.md section 4's refract / reflectance answers, then random inputs."""
    s = 1 / math.sqrt(2)
    refl, refr, f = gpu_ctx.eval_dielectric([[s, -s, 0], [0, -1, 0], [s, -s, 0]], [[0, 1, 0], [0, 1, 0], [0, 1, 0]], [1 / 1.5, 1 / 1.5, 1.5])
    assert np.allclose(refr[0], (0.4714045208, -0.8819171037, 0.0), atol=1e-6)
    assert np.allclose(refl[0], (s, s, 0.0), atol=1e-6)
    assert f[1] == pytest.approx(0.04, abs=1e-6)                       # reflectance(1.0, 1/1.5)
    lib = oracle.load()
    assert f[2] == pytest.approx(lib.oracle_reflectance(s, 1.5), abs=1e-6)
    rng = np.random.Generator(np.random.Philox(22))
    n = 100_000
    nn = rng.normal(size=(n, 3))
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    uv = rng.normal(size=(n, 3))
    uv /= np.linalg.norm(uv, axis=1, keepdims=True)
    uv = np.where((uv * nn).sum(1, keepdims=True) > 0, -uv, uv)  # incoming: against the normal
    ratio = np.where(rng.random(n) < 0.5, 1 / 1.5, 1.5)
    refl, refr, f = gpu_ctx.eval_dielectric(uv, nn, ratio)
    uv32, nn32, r32 = uv.astype(np.float32).astype(np.float64), nn.astype(np.float32).astype(np.float64), ratio.astype(np.float32).astype(np.float64)
    cos_t = np.minimum((-uv32 * nn32).sum(1), 1.0)
    perp = r32[:, None] * (uv32 + cos_t[:, None] * nn32)
    par = -np.sqrt(np.abs(1.0 - (perp * perp).sum(1)))[:, None] * nn32           # material.rs:114-119
    # |1 - perp^2| -> 0 at the critical angle: its square root magnifies the f32 rounding of perp there
    err, root = np.abs(refr - (perp + par)).max(axis=1), np.sqrt(np.abs(1.0 - (perp * perp).sum(1)))
    assert (err <= 1e-6 + 4e-7 / np.maximum(root, 1e-4)).all(), f"refract off by {err.max():.2e}"
    assert np.abs(refl - (uv32 - 2 * (uv32 * nn32).sum(1, keepdims=True) * nn32)).max() <= 1e-6
    r0 = ((1 - r32) / (1 + r32)) ** 2
    assert np.abs(f - (r0 + (1 - r0) * (1 - cos_t) ** 5)).max() <= 1e-6      # material.rs:121-125
    out = d3(0, 0, 0)
    for i in range(0, n, 997):  # the same numbers from the oracle's own functions
        lib.oracle_refract(d3(*uv32[i]), d3(*nn32[i]), r32[i], out)
        assert np.abs(refr[i] - np.array(list(out))).max() <= 1e-6 + 4e-7 / max(root[i], 1e-4)
        assert f[i] == pytest.approx(lib.oracle_reflectance(cos_t[i], r32[i]), abs=1e-6)


def test_perlin_noise_and_turbulence_on_device(rt, oracle, gpu_ctx):
    """perlin.rs:46-106 / texture.rs:57-65 on the device at 10^5 random points against the oracle."""
    api = rt.api
    hs = api.HostScene("two_perlin_spheres", seed=1)
    gsc = api.Scene(gpu_ctx, hs.desc)
    osc = oracle.OracleScene(hs.desc)
    desc = hs.desc.contents
    tab = desc.perlins[0]
    lib = oracle.load()
    rng = np.random.Generator(np.random.Philox(23))
    n = 100_000
    pts = np.concatenate([rng.uniform(-50, 50, (n // 2, 3)), rng.uniform(-1000, 1000, (n - n // 2, 3))])
    pts[:8] = np.round(pts[:8])  # lattice points: noise is exactly zero there
    noise, turb = gsc.eval_perlin(0, pts), gsc.eval_perlin(0, pts, turb_depth=7)
    want_n = np.array([lib.oracle_perlin_noise(C.byref(tab), d3(*p)) for p in pts])
    want_t = np.array([lib.oracle_perlin_turb(C.byref(tab), d3(*p), 7) for p in pts])
    assert np.abs(noise - want_n).max() <= 1e-5 and np.abs(noise[:8]).max() <= 1e-6
    assert np.abs(turb - want_t).max() <= 2e-5
    noise_tex = [i for i in range(desc.n_textures) if desc.textures[i].type == api.TEX_NOISE][0]
    got = gsc.eval_texture(noise_tex, pts[:20000])
    scale = desc.textures[noise_tex].scale
    want = 0.5 * (1 + np.sin(scale * pts[:20000, 2] + 10 * want_t[:20000]))
    assert np.abs(got - want[:, None]).max() <= 3e-4
    k = 7  # and through the oracle's own Texture::value for a few
    assert np.allclose(osc.texture_value(noise_tex, 0.0, 0.0, pts[k]), got[k], atol=3e-4)
    gsc.close()


def test_checker_and_image_textures_on_device(rt, oracle, gpu_ctx):
    """texture.rs:46-55 (checker on the hit position) and :67-89 (nearest texel, v flipped, texel / 255) on the device."""
    api = rt.api
    hs = api.HostScene("two_spheres", seed=1)
    gsc = api.Scene(gpu_ctx, hs.desc)
    desc = hs.desc.contents
    chk = [i for i in range(desc.n_textures) if desc.textures[i].type == api.TEX_CHECKER][0]
    assert np.allclose(gsc.eval_texture(chk, [[0.1, 0.1, 0.1], [0.1, 0.1, -0.1]]), [[0.2, 0.3, 0.1], [0.9, 0.9, 0.9]])  # main.rs:298-307
    rng = np.random.Generator(np.random.Philox(24))
    pts = rng.uniform(-30, 30, (100_000, 3))
    got = gsc.eval_texture(chk, pts)
    sines = np.sin(10 * pts[:, 0]) * np.sin(10 * pts[:, 1]) * np.sin(10 * pts[:, 2])
    want = np.where(sines[:, None] < 0, np.array([[0.9, 0.9, 0.9]]), np.array([[0.2, 0.3, 0.1]]))
    differ = np.abs(got - want).max(axis=1) > 1e-6
    assert differ.mean() <= 1e-4 and np.abs(sines[differ]).max(initial=0.0) < 1e-4  # only on a cell border within f32 rounding
    gsc.close()

    he = api.HostScene("earth", seed=1)
    gse = api.Scene(gpu_ctx, he.desc)
    de = he.desc.contents
    img = [i for i in range(de.n_textures) if de.textures[i].type == api.TEX_IMAGE][0]
    earth = api.load_earthmap()
    h, w, _ = earth.shape
    uv = np.concatenate([rng.uniform(-0.2, 1.2, (50_000, 2)), [[0.0, 1.0], [0.999999, 0.0], [0.5, 0.5], [1.5, -0.5], [0.25, 0.75]]]).astype(np.float32)
    got = gse.eval_texture(img, np.zeros((len(uv), 3)), uv)
    u = np.clip(uv[:, 0].astype(np.float64), 0, 1)
    v = 1.0 - np.clip(uv[:, 1].astype(np.float64), 0, 1)
    i, j = np.minimum((u * w).astype(np.int64), w - 1), np.minimum((v * h).astype(np.int64), h - 1)  # texture.rs:69-79
    want = earth[j, i] / 255.0
    differ = np.abs(got - want).max(axis=1) > 1e-6
    frac_u, frac_v = (u * w) % 1.0, (v * h) % 1.0
    on_border = (np.minimum(frac_u, 1 - frac_u) < 1e-3) | (np.minimum(frac_v, 1 - frac_v) < 1e-3)
    assert differ.mean() <= 2e-3 and on_border[differ].all()  # f32 u * W against f64 within a texel border
    assert not differ[-5:].any()
    gse.close()


def _onb_local(axis, x, y, z):
    """onb.rs:13-28."""
    w = axis / np.linalg.norm(axis, axis=1, keepdims=True)
    a = np.where(np.abs(w[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    v = np.cross(w, a)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    u = np.cross(w, v)
    return x[:, None] * u + y[:, None] * v + z[:, None] * w


def test_scatter_on_device(cornell, oracle):
    """`Material::scatter` + the mixture-pdf weight (main.rs:75-104, material.rs:70-161, pdf.rs:36-69) at the closest hits of
    the fixed ray set, path by path: the Philox draws are replayed on the host (rt1w_philox4x32), the formulas restated
    in f64 numpy, the light pdfs taken from the oracle."""
    api, hs, gsc, osc = cornell
    lib = oracle.load()
    n, seed = 1 << 16, 0x1234
    rays = make_ray_set(api, hs, osc, gsc.prims(), n)
    prim, mat, dirs, weight, time = gsc.eval_scatter(rays, seed=seed)
    op, ot, on, off, ouv, amb = osc.trace_closest(rays, seed=seed)
    keep = (amb == 0) & (op >= 0)
    assert (prim[amb == 0] == op[amb == 0]).all()
    prims = gsc.prims()
    desc = hs.desc.contents
    mat_id = np.array([prims[p].material if p >= 0 else -1 for p in op])
    mat_type = np.array([desc.materials[m].type if m >= 0 else -1 for m in mat_id])
    assert (mat[keep] == mat_type[keep]).all()
    o, d = rays["origin"].astype(np.float64), rays["direction"].astype(np.float64)
    p = o + np.where(keep, ot, 0.0)[:, None] * d
    x = np.array([api.philox4x32((i & 0xFFFF, 0, 1 ^ ((seed >> 32) << 4), 0), (i, seed & 0xFFFFFFFF)) if keep[i] else (0, 0, 0, 0) for i in range(n)],
                 dtype=np.uint64)
    u01 = (x >> 8).astype(np.float64) / 16777216.0
    unit = d / np.linalg.norm(d, axis=1, keepdims=True)
    dlen = np.linalg.norm(dirs.astype(np.float64), axis=1)

    # ---- Lambertian: direction from the mixture (pdf.rs:62-68), weight = albedo * scattering_pdf / pdf (main.rs:100-103)
    lam = keep & (mat_type == api.MAT_LAMBERTIAN)
    assert lam.sum() > n // 4
    r1, r2 = u01[:, 2], u01[:, 3]
    to_light = lam & ((x[:, 0] >> 31) == 1)
    pick = np.minimum((u01[:, 1] * 2).astype(np.int64), 1)  # slice.choose over the two lights (hittable.rs:153)
    rect = to_light & (pick == 0)
    want = np.stack([LIGHT[0] + (LIGHT[1] - LIGHT[0]) * r1 - p[:, 0], LIGHT[4] - p[:, 1], LIGHT[2] + (LIGHT[3] - LIGHT[2]) * r2 - p[:, 2]], axis=1)
    assert rect.sum() > 1000 and np.abs(dirs[rect] - want[rect]).max() <= 1e-3          # aarect.rs:140-147, un-normalised
    sph = to_light & (pick == 1)
    axis = np.where(sph[:, None], np.array([GLASS[0]]) - p, on)
    with np.errstate(divide="ignore", invalid="ignore"):  # rows of rays that missed (axis = 0) are masked out below
        cos_max = np.sqrt(np.maximum(1 - GLASS[1] ** 2 / (axis * axis).sum(1), 0.0))
        z = np.where(sph, 1 + r2 * (cos_max - 1), np.sqrt(1 - r2))                          # math.rs:39-65
        q = np.sqrt(np.maximum(1 - z * z, 0.0))
        local = _onb_local(axis, np.cos(2 * np.pi * r1) * q, np.sin(2 * np.pi * r1) * q, z)
    onb = lam & ~rect
    assert sph.sum() > 1000 and np.abs(dirs[onb] - local[onb]).max() <= 2e-4
    nrm = on / np.maximum(np.linalg.norm(on, axis=1, keepdims=True), 1e-30)
    dd = dirs.astype(np.float64)
    cosine = np.maximum((dd / np.maximum(dlen, 1e-30)[:, None] * nrm).sum(1) / np.pi, 0.0)
    idx = np.flatnonzero(lam)
    light_pdf = np.zeros(n)
    rect5, c3 = d3(*LIGHT), d3(*GLASS[0])
    for i in idx:
        oi, vi = d3(*p[i]), d3(*dd[i])
        light_pdf[i] = 0.5 * (lib.oracle_xz_rect_pdf_value(rect5, oi, vi) + lib.oracle_sphere_pdf_value(c3, GLASS[1], oi, vi))
    pdf = 0.5 * light_pdf + 0.5 * cosine
    albedo = np.array([list(desc.textures[desc.materials[m].texture].color) if m >= 0 and desc.materials[m].texture >= 0 else [0, 0, 0] for m in mat_id])
    want_w = albedo * (cosine / np.maximum(pdf, 1e-300))[:, None]
    sane = lam & (pdf > 1e-6) & np.isfinite(light_pdf)
    rel = np.abs(weight[sane] - want_w[sane]).max(axis=1) / np.maximum(want_w[sane].max(axis=1), 1e-3)
    assert np.quantile(rel, 0.995) <= 1e-3, f"lambertian weight off by {np.quantile(rel, 0.995):.2e} (99.5 % quantile)"
    assert np.abs(time[lam] - ot[lam]).max() <= 1e-5 * ot[lam].max()                       # main.rs:86: time = hit t

    # ---- Metal (the tall box, fuzz 0): reflect(unit(d), n), attenuation = albedo, ray.time kept (material.rs:98-112)
    met = keep & (mat_type == api.MAT_METAL)
    assert met.sum() > 1000
    want = unit - 2 * (unit * on).sum(1, keepdims=True) * on
    assert np.abs(dirs[met] - want[met]).max() <= 1e-5
    m_alb = np.array([list(desc.materials[m].albedo) if m >= 0 else [0, 0, 0] for m in mat_id])
    assert np.abs(weight[met] - m_alb[met]).max() <= 1e-6 and np.array_equal(time[met], rays["time"][met])

    # ---- Dielectric (the glass sphere): Schlick against the path's first draw (material.rs:132-161)
    die = keep & (mat_type == api.MAT_DIELECTRIC)
    assert die.sum() > 1000
    ratio = np.where(off != 0, 1 / 1.5, 1.5)
    cos_t = np.minimum((-unit * on).sum(1), 1.0)
    sin_t = np.sqrt(np.maximum(1 - cos_t * cos_t, 0.0))
    r0 = ((1 - ratio) / (1 + ratio)) ** 2
    schlick = r0 + (1 - r0) * (1 - cos_t) ** 5
    reflects = (ratio * sin_t > 1.0) | (schlick > u01[:, 0])
    perp = ratio[:, None] * (unit + cos_t[:, None] * on)
    refr = perp - np.sqrt(np.abs(1 - (perp * perp).sum(1)))[:, None] * on
    want = np.where(reflects[:, None], unit - 2 * (unit * on).sum(1, keepdims=True) * on, refr)
    sure = die & (np.abs(ratio * sin_t - 1.0) > 1e-4) & (np.abs(schlick - u01[:, 0]) > 1e-4)
    assert np.abs(dirs[sure] - want[sure]).max() <= 2e-5
    assert np.array_equal(weight[die], np.ones((die.sum(), 3), np.float32)) and np.array_equal(time[die], rays["time"][die])

    # ---- DiffuseLight: emits on the front face only, and FlipFace has toggled which that is (material.rs:168-181, main.rs:470-477)
    lit = keep & (mat_type == api.MAT_DIFFUSE_LIGHT)
    assert lit.sum() > 100
    assert np.array_equal(weight[lit], np.where(off[lit, None] != 0, np.float32(15.0), np.float32(0.0)) * np.ones((1, 3), np.float32))
    assert (dlen[lit] == 0).all()
