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
