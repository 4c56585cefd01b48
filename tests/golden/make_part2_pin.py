"""Generates tests/golden/part2_final_pin.npz from the reference's committed render part2_final.png (600 x 800).

Run in the build container only (/root/reference is not on the GPU box).  part2_final.png was rendered by
examples/part2_all.rs: camera (-9,3,-9) -> (1,3,2), fov 25, 600x800 (part2_all.rs:88-98).  Two parts of that scene come
from tiny_rng::Rng::new(12345) (crate not vendored): the heights of the 400 floor boxes and the positions of the 1000
small spheres.  Everything else is fixed by the example's constants: the light (XZRect), the brown Lambertian sphere, the
TurbulenceTexture(5, 10) sphere, the earth ImageTexture sphere, the glass sphere holding a blue ConstantMedium, the
MetalMat(roughness 10) sphere and the global fog medium.  The fixture keeps the PNG's 8x8 box means and a mask of the
boxes whose 64 primary rays all hit one of those fixed objects (first-hit ids from the CPU oracle on the scene
scenes/part2_all.yml), so the GPU render can be compared with the reference's own output exactly where the scene is
reproducible: TurbulenceTexture / Perlin (texture.rs:113-224), ImageTexture uv on a sphere (sphere.rs:22-29,
texture.rs:296-309), DielectricMat (material.rs:122-150), ConstantMedium + IsotropicMat (volume.rs:57-82,
material.rs:198-203), MetalMat with roughness (material.rs:91-106), EmissiveMat, Camera.
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import oracle_scene, params_for  # noqa: E402

W, H, K = 600, 800, 8
img = np.asarray(Image.open("/root/reference/part2_final.png").convert("RGB")).astype(np.float64)
assert img.shape == (H, W, 3)
low = img.reshape(H // K, K, W // K, K, 3).mean((1, 3))

orc = oracle_scene("part2_all", fast=True)
p = params_for("part2_all", W, H, 4, seed=1)
FIXED = {"light": 400, "brown": 401, "glass_small": 402, "metal": 403, "glass": 404, "medium": 405, "earth": 406, "noise": 407}
votes = np.zeros((H, W), np.int32)
first = np.full((H, W), -1, np.int32)
for s in range(4):
    o, d = orc.primary_rays(p, s)
    obj = orc.first_hit(o, d, seed=1, pixel=np.arange(W * H, dtype=np.uint32), sample=np.full(W * H, s, np.uint32))["obj"].reshape(H, W)
    votes += (obj >= 400) & (obj <= 407)
    first = np.where(first < 0, obj, first)
inside = (votes == 4).reshape(H // K, K, W // K, K).all((1, 3))                 # every jittered primary ray of the box hits a fixed object
label = np.zeros((H // K, W // K), np.int8)
for name, idx in FIXED.items():
    m = (first == idx).reshape(H // K, K, W // K, K).mean((1, 3)) > 0.9
    label[m & inside] = idx - 399
np.savez_compressed(os.path.join(HERE, "part2_final_pin.npz"), low=np.round(low, 2).astype(np.float32), mask=inside, label=label,
                    size=np.array([W, H], np.int32))
print("boxes in the mask:", int(inside.sum()), "of", inside.size, {n: int((label == i - 399).sum()) for n, i in FIXED.items()})
