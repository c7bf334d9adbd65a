"""The image output path (main.rs:953,992,1003-1007; color.rs:14-21,56-65): device resolve and the demo driver."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_resolve_is_bit_exact(rt, gpu_ctx):
    """rt1w_render_rgb8 (resolve on the device) == rt1w_resolve_rgb8 (host) of the same sums.
    One sample per pixel: every pixel receives a single atomic add per channel, so the two renders are identical."""
    api = rt.api
    for name in ("cornel_box", "random_scene"):
        hs = api.HostScene(name, seed=1)
        gsc = api.Scene(gpu_ctx, hs.desc)
        cam = hs.camera()
        p = hs.params(width=160, spp=1, seed=11)
        rgb_sum, _, _ = gsc.render(cam, p)
        rgb8, st = gsc.render_rgb8(cam, p)
        assert st.paths == p.width * p.height
        assert rgb8.shape == (p.height, p.width, 3) and rgb8.dtype == np.uint8
        assert np.array_equal(rgb8, api.resolve_rgb8(rgb_sum, 1))
        assert rgb8.max() > 0
        gsc.close()


def test_resolve_rules(rt):
    """NaN sum -> 0, gamma 2, clamp at 0.999 * 256 (color.rs:14-21,56-65) - host resolve used as the checker above."""
    api = rt.api
    x = np.array([[[np.nan, 0.25, 4.0], [0.0, 1.0, 1e-6]]], dtype=np.float32)
    out = api.resolve_rgb8(x, 1)
    assert out.tolist() == [[[0, 128, 255], [0, 255, 0]]]


def test_demo_driver_writes_the_reference_ppm(rt, gpu_ctx):
    """rt1w_main prints `P3\\n{w} {h}\\n255\\n` and one `r g b` line per pixel, top row first (main.rs:953,1003-1007)."""
    api = rt.api
    exe = os.path.join(os.path.dirname(api.LIB_PATH), "rt1w_main")
    r = subprocess.run([exe, "cornel_box", "--width", "40", "--spp", "64"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Scanlines remaining" in r.stderr and r.stderr.rstrip().endswith("Done")
    lines = r.stdout.split("\n")
    assert lines[0] == "P3" and lines[1] == "40 40" and lines[2] == "255"
    body = np.array([[int(v) for v in ln.split()] for ln in lines[3:] if ln], dtype=np.int64)
    assert body.shape == (40 * 40, 3) and body.min() >= 0 and body.max() <= 255
    # the same render through the Python binding: same Philox streams, fp32 summation order may differ
    hs = api.HostScene("cornel_box", seed=1)
    gsc = api.Scene(gpu_ctx, hs.desc)
    rgb8, _ = gsc.render_rgb8(hs.camera(), hs.params(width=40, spp=64, seed=0))
    diff = np.abs(body.reshape(40, 40, 3) - rgb8.astype(np.int64))
    assert (diff <= 1).mean() >= 0.999
    gsc.close()
