"""The native command-line driver (csrc/cli_main.cpp, the counterpart of src/main.rs) against the Python one: same scene file,
same hard-coded camera, same samples -> the same PNG, byte for byte in pixel values."""
import os
import subprocess

import numpy as np
import pytest

from conftest import CONFIGS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["conics", "suzanne"])
def test_native_cli_renders_the_same_image_as_the_python_cli(tmp_path, name):
    from PIL import Image
    from firework_b200.__main__ import main
    from firework_b200.assets import load_image_native
    from firework_b200.build import CLI
    scene = CONFIGS[name].path()
    out_native, out_py = str(tmp_path / "native.png"), str(tmp_path / "py.png")
    r = subprocess.run([CLI, "--scene-file", scene, "-s", "6", "--seed", "11", "-o", out_native, "-n", "x"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Finished Rendering in" in r.stdout and f'Saving image to "{out_native}"' in r.stdout
    assert main(["--scene-file", scene, "-s", "6", "--seed", "11", "-o", out_py]) == 0
    a = np.asarray(Image.open(out_native).convert("RGB"))
    b = np.asarray(Image.open(out_py).convert("RGB"))
    assert a.shape == (540, 960, 3)                      # main.rs:33-34
    assert np.array_equal(a, b)
    assert np.array_equal(load_image_native(out_native)[..., :3], a)
    assert a.std() > 5                                   # an image, not a constant


def test_native_cli_default_output_name_and_multi_gpu(tmp_path):
    from PIL import Image
    from firework_b200 import _native as N
    from firework_b200.build import CLI
    scene = CONFIGS["conics"].path()
    r = subprocess.run([CLI, "--scene-file", scene, "--samples=4", "--width", "240", "--height", "135", "-n", "My Render"],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    one = np.asarray(Image.open(tmp_path / "My_Render.png").convert("RGB")).astype(int)
    if N.lib().fw_device_count() >= 2:
        r = subprocess.run([CLI, "--scene-file", scene, "-s", "4", "--width", "240", "--height", "135", "--gpus", "2", "-o", str(tmp_path / "two.png")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        two = np.asarray(Image.open(tmp_path / "two.png").convert("RGB")).astype(int)
        assert np.abs(one - two).max() <= 1              # fp32 order of the cross-GPU sum


def test_cpp_examples_render_what_the_python_mirror_renders(tmp_path):
    """examples/cornell_box.cpp and examples/earth.cpp through include/firework.hpp -> the C ABI, against the Python mirror of the
    same example with the same renderer settings: identical pixels (earth: with the natively decoded texels on both sides)."""
    from PIL import Image
    from test_host import _build_example
    from firework_b200 import scenes
    from firework_b200.assets import load_image_native
    from firework_b200.scenes import SCENE_DIR
    assets = os.path.join(SCENE_DIR, "assets")
    exe = _build_example("cornell_box", tmp_path)
    out = str(tmp_path / "cb.png")
    r = subprocess.run([exe, "-s", "16", "--width", "72", "--height", "64", "--seed", "5", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = CONFIGS["cornell_box"].renderer(width=72, height=64, samples=16, seed=5).render(scenes.cornell_box())
    assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), want)

    exe = _build_example("earth", tmp_path)
    out = str(tmp_path / "earth.png")
    r = subprocess.run([exe, "-s", "8", "--width", "96", "--height", "96", "--seed", "2", "--asset-dir", assets, "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sc = scenes.earth_scene()
    for f in ("earthmap.jpg", "uvmap.png"):
        sc.register_asset(f, load_image_native(os.path.join(assets, f)))
    want = CONFIGS["earth"].renderer(width=96, height=96, samples=8, seed=2).render(sc)
    assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), want)
    r = subprocess.run([exe, "-s", "1", "--asset-dir", str(tmp_path / "nowhere")], capture_output=True, text=True)
    assert r.returncode == 1 and "not found" in r.stderr      # errors are messages, not aborts
