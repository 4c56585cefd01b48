"""Generates tests/golden/reference_png_lowres.npz from the renders committed in the reference repository.

Run in the build container only (/root/reference is not on the GPU box).  These PNGs are renders of scenes that can be
reproduced exactly — the serde dumps scenes/suzanne.yml, teapot.yml, conics.yml, and the deterministic scene
constructors of examples/cornell_box.rs, earth.rs, heightmap.rs — with cameras fixed in the examples (unlike
random_spheres.png / part2_final.png, whose scenes come from an RNG crate that is not vendored).  The fixture holds
their 8x8 box-filtered RGB means (noise-suppressed, 1/64 of the pixels); the GPU test renders the same scenes at the
same resolution and compares box means (tests/test_gpu_parity.py).
"""
import os
import numpy as np
from PIL import Image

REF = "/root/reference"
# volume.png was rendered by an earlier revision of examples/volume_test.rs (an r = 1.5 medium sphere in a concentric glass
# shell; geometry, light reflection and horizon fit that, colour / density do not: tools notes in profiles/r02_variants.md §5),
# so its sphere cannot be reproduced — but its background (sky, horizon, far floor: everything the sphere does not influence)
# is the current example's, and the test compares only those boxes.
FILES = {"suzanne": "suzanne.png", "teapot": "teapot.png", "cornell_box": "cornell_box.png", "conics": "conics.png",
         "earth": "Earth.png", "heightmap": "heightmap.png", "volume": "volume.png"}
out = {}
for name, fn in FILES.items():
    img = np.asarray(Image.open(os.path.join(REF, fn)).convert("RGB")).astype(np.float64)
    h, w, _ = img.shape
    low = img[:h // 8 * 8, :w // 8 * 8].reshape(h // 8, 8, w // 8, 8, 3).mean((1, 3))
    out[name] = np.round(low, 2).astype(np.float32)
    out[name + "_size"] = np.array([w, h], np.int32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_png_lowres.npz"), **out)
print({k: v.shape for k, v in out.items()})
