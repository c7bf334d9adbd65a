"""Decodes the reference's assets/earthmap.jpg (main.rs:347,748) into assets/earthmap.ppm (binary P6).

The reference decodes the JPEG with the `image` crate at start-up; no C/C++ JPEG decoder exists in
this image, so decoding is a one-off asset-prep step done with PIL.  Different decoders may differ by
a level or two per texel; the product and the oracle both read this PPM.

Run once in the authoring container:  python tools/prep_earthmap.py [/root/reference/assets/earthmap.jpg]
"""
import os
import sys

from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/earthmap.jpg"
    im = Image.open(src).convert("RGB")
    dst = os.path.join(ROOT, "assets", "earthmap.ppm")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % im.size)
        f.write(im.tobytes())
    print(dst, im.size)


if __name__ == "__main__":
    main()
