"""Decode the reference's JPEG asset once to a binary PPM next to it.

The reference decodes `earthmap.jpg` with the `image` crate on the host (src/texture.rs:79).  The C++
host mirror reads the decoded sibling `earthmap.ppm`; JPEG entropy decoding is not part of the hot
path and is done here with PIL.  Run as `python -m grayshift_b200.assets`.
"""
import os

from .ffi import ASSET_DIR


def ensure_decoded(asset_dir=ASSET_DIR):
    src = os.path.join(asset_dir, "earthmap.jpg")
    dst = os.path.join(asset_dir, "earthmap.ppm")
    if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
        return dst
    from PIL import Image

    Image.open(src).convert("RGB").save(dst)
    return dst


if __name__ == "__main__":
    print(ensure_decoded())
