"""Writes tests/golden/textures/lines.png: a stand-in for the `textures/lines.png` the reference's
src/scenes/spheres-differentials-texfilt.pbrt names but does not ship.  96 x 64 (NOT a power of two: the MIPMap
resampling path runs), thin bright lines on a dark ground plus a colour ramp, so that texture filtering has something to do."""
from pathlib import Path

import numpy as np
from PIL import Image

h, w = 64, 96
y, x = np.mgrid[0:h, 0:w]
img = np.zeros((h, w, 3), np.uint8)
img[..., 0] = 30 + (x * 2) % 90
img[..., 1] = 40 + (y * 3) % 70
img[..., 2] = 60
img[(x % 12) < 2] = (250, 250, 240)
img[(y % 16) < 2] = (240, 60, 50)
out = Path(__file__).resolve().parent / "textures" / "lines.png"
Image.fromarray(img).save(out)
print("wrote", out)
