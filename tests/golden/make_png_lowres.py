"""Generates tests/golden/reference_png_lowres.npz from the renders committed in the reference repository.

Run in the build container only (/root/reference is not on the GPU box).  suzanne.png (960x540, examples/suzanne.rs)
and teapot.png (1920x1080, examples/teapot.rs) are renders of scenes whose serde dumps are committed next to them
(scenes/suzanne.yml, scenes/teapot.yml) with cameras fixed in the examples, so — unlike random_spheres.png — they can
be reproduced.  The fixture holds their 8x8 box-filtered RGB means (noise-suppressed, 1/64 of the pixels); the GPU test
renders the same scenes at the same resolution and compares box means (tests/test_gpu_parity.py).
"""
import os
import numpy as np
from PIL import Image

REF = "/root/reference"
out = {}
for name in ("suzanne", "teapot"):
    img = np.asarray(Image.open(os.path.join(REF, f"{name}.png")).convert("RGB")).astype(np.float64)
    h, w, _ = img.shape
    low = img[:h // 8 * 8, :w // 8 * 8].reshape(h // 8, 8, w // 8, 8, 3).mean((1, 3))
    out[name] = np.round(low, 2).astype(np.float32)
    out[name + "_size"] = np.array([w, h], np.int32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_png_lowres.npz"), **out)
print({k: v.shape for k, v in out.items()})
