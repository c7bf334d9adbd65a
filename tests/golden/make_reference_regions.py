"""Region statistics of the reference's published render, the only end-to-end result pin it ships.

/root/reference/rest_of_your_life.png is the output of master's main.rs scene 5 (Cornell box, 600x600,
100 spp; README.md:19-21).  Run in the authoring container (the GPU box has no /root/reference):

    python tests/golden/make_reference_regions.py     # writes tests/golden/rest_of_your_life_regions.json

Pixel values are inverted through Display for SampledColor (color.rs:56-65): mean radiance ~ (v / 256)^2.
"""
import json
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REGIONS = {  # name: (row0, row1, col0, col1), row 0 = top
    "whole": (0, 600, 0, 600),
    "back_wall": (150, 250, 200, 400),
    "green_wall_image_left": (200, 400, 30, 110),
    "red_wall_image_right": (200, 400, 490, 570),
    "floor_front": (540, 575, 150, 450),
    "ceiling": (30, 70, 150, 450),
    "box_front_face": (280, 440, 200, 250),
    "light": (85, 95, 260, 340),
}


def main():
    im = np.asarray(Image.open("/root/reference/rest_of_your_life.png").convert("RGB")).astype(np.float64)
    assert im.shape == (600, 600, 3)
    lin = (im / 256.0) ** 2
    out = {"source": "rest_of_your_life.png (600x600, scene 5, 100 spp)", "regions": {}}
    for name, (r0, r1, c0, c1) in REGIONS.items():
        out["regions"][name] = {"rows": [r0, r1], "cols": [c0, c1], "mean_rgb8": im[r0:r1, c0:c1].mean(axis=(0, 1)).round(3).tolist(),
                                "mean_linear": lin[r0:r1, c0:c1].mean(axis=(0, 1)).round(6).tolist()}
    nonblack = np.flatnonzero(im.sum(axis=2).max(axis=1) > 0)
    out["first_last_nonblack_row"] = [int(nonblack[0]), int(nonblack[-1])]
    nonblack_c = np.flatnonzero(im.sum(axis=2).max(axis=0) > 0)
    out["first_last_nonblack_col"] = [int(nonblack_c[0]), int(nonblack_c[-1])]
    with open(os.path.join(HERE, "rest_of_your_life_regions.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
