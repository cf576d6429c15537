"""Writes tests/golden/letterbox.npz: outputs of the REFERENCE's resize_image (utils/utils.py:16-30, i.e. Pillow
Image.resize BICUBIC + grey canvas) on small seeded images.  Runs only in the authoring container (/root/reference and
Pillow present); the fixture pins oracle/letterbox_oracle.py and the CUDA kernel.  Pillow version is stored alongside.

usage: python -m oracle.make_golden_letterbox
"""
import os

import numpy as np

from oracle import ref_import

# (source w, source h, canvas w, canvas h): up- and down-scaling, both letterbox orientations, identity, extreme ratios
CASES = [(97, 61, 96, 64), (61, 97, 96, 64), (50, 38, 64, 64), (64, 64, 64, 64), (200, 33, 64, 64), (7, 5, 32, 32),
         (160, 120, 64, 48), (300, 200, 32, 32), (48, 64, 64, 64)]


def main():
    ref_import.import_reference()
    import PIL
    from PIL import Image
    import utils.utils as RU                      # the reference module itself
    rng = np.random.default_rng(2024)
    out = {"pillow_version": np.array(PIL.__version__), "cases": np.array(CASES, np.int32)}
    for i, (iw, ih, w, h) in enumerate(CASES):
        img = rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        if i % 3 == 0:                            # smooth content as well as noise
            yy, xx = np.mgrid[0:ih, 0:iw]
            img = np.stack([(xx * 255 // max(iw - 1, 1)), (yy * 255 // max(ih - 1, 1)), ((xx + yy) % 256)], -1).astype(np.uint8)
        out[f"src{i}"] = img
        for lb in (0, 1):
            out[f"dst{i}_lb{lb}"] = np.asarray(RU.resize_image(Image.fromarray(img), (w, h), bool(lb)))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "letterbox.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
