"""Generates tests/golden/checkerboard_{square,stretched}.npz: the decoded RGB8 texels of the
reference's texture fixtures (src/texture/testdata/*.png), so that the imagemap.rs known-answer
tests can run where /root/reference does not exist.  Decoded with Pillow AND with the package's
own PNG reader; both must agree."""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrt_rust_b200.imageio import read_png_rgb8  # noqa: E402

SRC = "/root/reference/src/texture/testdata"
for name in ("checkerboard_square", "checkerboard_stretched"):
    path = os.path.join(SRC, name + ".png")
    ref = np.asarray(Image.open(path).convert("RGB"), np.uint8)
    own = read_png_rgb8(path)
    assert ref.shape == own.shape and np.array_equal(ref, own), name
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), rgb8=own)
    print(name, own.shape, "unique values", np.unique(own).tolist()[:8])
