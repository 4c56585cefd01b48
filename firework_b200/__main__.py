"""`python -m firework_b200 --scene-file X.yml -s N [-o out.png] [-n name]` — the reference's CLI
(src/main.rs:6-62) on the GPU path: same flags, same hard-coded camera ((0,30,50) -> origin, fov 40), 960x540,
BVH on.  `-o` writes a PNG (window.rs:59-66 `save_image`); without it the image is written to `<name>.png`
(there is no window here).  Extra flags: --checkpoint FILE resumes / saves the fp32 sample sums so that a long
render can be interrupted (the reference's 3.36 h render has no intermediate save, README.md:16)."""
import argparse
import os
import sys
import time

import numpy as np

from . import CameraSettings, Renderer, Scene
from .progressive import render_progressive


def main(argv=None):
    ap = argparse.ArgumentParser(prog="firework")
    ap.add_argument("--scene-file", required=True)
    ap.add_argument("-n", "--name", default=None)
    ap.add_argument("-s", "--samples", type=int, required=True)
    ap.add_argument("-o", "--output", default=None)
    ap.add_argument("--checkpoint", default=None, help="npz file holding the fp32 sums + samples done (resume / save)")
    ap.add_argument("--chunk", type=int, default=0, help="samples per progressive step (default: all at once)")
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--seed", type=int, default=0)
    opt = ap.parse_args(argv)

    scene = Scene.from_file(opt.scene_file)
    camera = CameraSettings.default().cam_pos((0.0, 30.0, 50.0)).look_at((0.0, 0.0, 0.0)).field_of_view(40.0)
    renderer = (Renderer.default().width(opt.width).height(opt.height).samples(opt.samples).use_bvh(True)
                .camera(camera).seed(opt.seed))
    start = time.time()
    rgb, info = render_progressive(scene, renderer, checkpoint=opt.checkpoint, chunk=opt.chunk or opt.samples)
    print(f"Finished Rendering in {int(time.time() - start)} s")
    name = opt.name or "Firework Render"
    out = opt.output or (name.replace(" ", "_") + ".png")
    print(f"Saving image to {out!r}")
    from PIL import Image
    Image.fromarray(rgb).save(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
