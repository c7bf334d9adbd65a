"""Noise of a 600x600, 100-spp Cornell render, region by region: what |oracle - reference PNG| may be.

The reference's published render (rest_of_your_life.png) is ONE 100-spp realisation; so is an oracle render of the same
configuration.  This script renders the oracle eight times (sample ranges [100 k, 100 (k + 1)), k = 0..7: independent
per-pixel RNG streams), pushes every image through the file's own quantisation (color.rs:14-21,56-65, then (v / 256)^2),
and writes per-region standard deviations of the region means - plus the spread of the other pins (share of pure-black
pixels inside the frame, share of saturated pixels on the light) - into tests/golden/rest_of_your_life_regions.json
next to the PNG's own values.  Two independent realisations differ by sqrt(2) sigma; the tests allow 4 sqrt(2) sigma.

    python tests/golden/make_reference_regions.py && python tests/golden/make_oracle_sigma.py      (about 8 CPU-minutes)
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def linear_image(rgb_sum, spp):
    x = np.where(np.isnan(rgb_sum), 0.0, rgb_sum) / spp
    q = np.floor(256.0 * np.clip(np.sqrt(np.maximum(x, 0.0)), 0.0, 0.999))
    return (q / 256.0) ** 2


def frame_stats(lin, golden):
    """Share of pure-black pixels inside the non-black frame, share of saturated (255) values on the light."""
    r0, r1 = golden["first_last_nonblack_row"]
    c0, c1 = golden["first_last_nonblack_col"]
    inner = lin[r0:r1 + 1, c0:c1 + 1]
    g = golden["regions"]["light"]
    light = lin[g["rows"][0]:g["rows"][1], g["cols"][0]:g["cols"][1]]
    return float((inner.max(axis=2) == 0.0).mean()), float((light >= (255.0 / 256.0) ** 2).mean())


def main():
    import oracle_binding

    api = importlib.import_module("raytracing-1w_b200").api
    path = os.path.join(HERE, "rest_of_your_life_regions.json")
    golden = json.load(open(path))
    if os.path.exists("/root/reference/rest_of_your_life.png"):
        from PIL import Image

        im = np.asarray(Image.open("/root/reference/rest_of_your_life.png").convert("RGB")).astype(np.float64)
        black, sat = frame_stats((im / 256.0) ** 2, golden)
        golden["black_fraction_inside_frame"], golden["light_saturated_fraction"] = black, sat
    hs = api.HostScene("cornel_box", seed=1)
    osc = oracle_binding.OracleScene(hs.desc)
    cam = hs.camera()
    means = {name: [] for name in golden["regions"]}
    blacks, sats = [], []
    for k in range(8):
        img, _, _ = osc.render(cam, hs.params(width=600, spp=100 * (k + 1), sample_begin=100 * k))
        lin = linear_image(img, 100)
        for name, g in golden["regions"].items():
            means[name].append(lin[g["rows"][0]:g["rows"][1], g["cols"][0]:g["cols"][1]].mean(axis=(0, 1)))
        b, s = frame_stats(lin, golden)
        blacks.append(b), sats.append(s)
        print(f"render {k}: whole-image mean {means['whole'][-1]}, black inside frame {b:.4f}, light saturated {s:.4f}", flush=True)
    for name, g in golden["regions"].items():
        m = np.array(means[name])
        g["oracle_mean_linear"] = m.mean(axis=0).round(6).tolist()
        g["oracle_sigma_linear"] = m.std(axis=0, ddof=1).round(7).tolist()
    golden["oracle_black_fraction"] = [float(np.mean(blacks)), float(np.std(blacks, ddof=1))]
    golden["oracle_light_saturated_fraction"] = [float(np.mean(sats)), float(np.std(sats, ddof=1))]
    golden["sigma_source"] = "8 oracle renders of 600x600 x 100 spp (tests/golden/make_oracle_sigma.py)"
    with open(path, "w") as f:
        json.dump(golden, f, indent=1)


if __name__ == "__main__":
    main()
