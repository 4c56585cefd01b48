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


def test_checkpoints_move_between_the_native_and_the_python_driver(tmp_path):
    """--checkpoint x.fwck: one file format for both drivers (cli_main.cpp / progressive.py).  A render stopped after 4 of 10
    samples by one driver and finished by the other is the 10-sample render (up to the fp32 order of adding the chunks)."""
    from PIL import Image
    from firework_b200.__main__ import main
    from firework_b200.build import CLI
    from firework_b200.progressive import fwck_load
    scene = CONFIGS["conics"].path()
    size = ["--width", "240", "--height", "135", "--seed", "7"]

    def img(p):
        return np.asarray(Image.open(p).convert("RGB")).astype(int)

    whole = str(tmp_path / "whole.png")
    assert subprocess.run([CLI, "--scene-file", scene, "-s", "10", "-o", whole] + size, capture_output=True, text=True).returncode == 0
    # native alone, in chunks
    ck = str(tmp_path / "a.fwck")
    r = subprocess.run([CLI, "--scene-file", scene, "-s", "10", "--chunk", "4", "--checkpoint", ck, "-o", str(tmp_path / "a.png")] + size,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sums, done, seed, _ = fwck_load(ck)
    assert done == 10 and seed == 7 and sums.shape == (135, 240, 3)
    assert np.abs(img(tmp_path / "a.png") - img(whole)).max() <= 1
    # native starts, python finishes
    ck = str(tmp_path / "b.fwck")
    assert subprocess.run([CLI, "--scene-file", scene, "-s", "4", "--checkpoint", ck, "-o", str(tmp_path / "b4.png")] + size,
                          capture_output=True, text=True).returncode == 0
    assert fwck_load(ck)[1] == 4
    assert main(["--scene-file", scene, "-s", "10", "--chunk", "3", "--checkpoint", ck, "-o", str(tmp_path / "b.png")] + size) == 0
    assert fwck_load(ck)[1] == 10
    assert np.abs(img(tmp_path / "b.png") - img(whole)).max() <= 1
    # python starts, native finishes
    ck = str(tmp_path / "c.fwck")
    assert main(["--scene-file", scene, "-s", "4", "--checkpoint", ck, "-o", str(tmp_path / "c4.png")] + size) == 0
    r = subprocess.run([CLI, "--scene-file", scene, "-s", "10", "--chunk", "5", "--checkpoint", ck, "-o", str(tmp_path / "c.png")] + size,
                       capture_output=True, text=True)
    assert r.returncode == 0 and "Resuming at sample 4 of 10" in r.stdout, r.stderr
    assert np.abs(img(tmp_path / "c.png") - img(whole)).max() <= 1
    # a checkpoint of another render is refused
    r = subprocess.run([CLI, "--scene-file", scene, "-s", "10", "--checkpoint", ck, "-o", str(tmp_path / "d.png"), "--width", "240", "--height", "135", "--seed", "8"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "does not match" in r.stderr


def test_cpp_mesh_example_renders_the_committed_scene(tmp_path):
    """examples/suzanne.cpp builds its scene from suzanne.obj through the C++ mirror; the reference's committed scenes/suzanne.yml is
    the same document (tests/test_host.py), so both must render the same pixels."""
    from PIL import Image
    from conftest import native_scene, params_for
    from test_host import _build_example
    from firework_b200.scenes import SCENE_DIR
    exe = _build_example("suzanne", tmp_path)
    out = str(tmp_path / "suzanne.png")
    r = subprocess.run([exe, "--obj", os.path.join(SCENE_DIR, "assets", "suzanne.obj"), "-s", "6", "--width", "160", "--height", "90", "--seed", "3", "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ns = native_scene("suzanne")
    p = params_for("suzanne", 160, 90, 6, seed=3)        # the example's camera (suzanne.rs:88-91), use_bvh(true)
    want, _, _ = ns.render(p, want_sum=False)
    ns.close()
    assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), want)
