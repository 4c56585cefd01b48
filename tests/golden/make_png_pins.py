"""Generates tests/golden/reference_png_pins.json from the PNG renders committed in the reference repository.

Run in the build container only (/root/reference is not on the GPU box).  The reference has no tests or golden
vectors; its committed renders are the only artefacts that pin behaviour.  cornell_box.png (300x300, produced
by examples/cornell_box.rs) is used numerically: mean u8 colour of seven flat patches.  The other images
(random_spheres.png etc.) were rendered with older camera settings / a scene RNG that cannot be reproduced, so
they pin orientation only (checked by eye).
"""
import json, os
import numpy as np
from PIL import Image

REF = "/root/reference"
PATCHES = {  # name: (y0, y1, x0, x1) in the 300x300 image
    "tall_box_front": (150, 230, 100, 145), "short_box_front": (215, 270, 155, 215), "short_box_top": (198, 204, 160, 215),
    "green_wall": (100, 200, 15, 45), "red_wall": (100, 200, 255, 285), "back_wall": (90, 120, 120, 200),
    "floor": (275, 290, 100, 200),
}
img = np.asarray(Image.open(os.path.join(REF, "cornell_box.png")).convert("RGB")).astype(np.float64)
out = {"source": "cornell_box.png (reference repo root), 300x300", "patches": {}}
for k, (y0, y1, x0, x1) in PATCHES.items():
    out["patches"][k] = {"box": [y0, y1, x0, x1], "mean_rgb": [round(float(v), 3) for v in img[y0:y1, x0:x1].mean((0, 1))]}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_png_pins.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
