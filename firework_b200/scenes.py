"""The benchmark / parity configurations of BASELINE.json, restated as data through the mirrored API.

Each builder follows the reference example line by line (cited), with two stated substitutions:
  * scene-generation randomness: the examples use tiny_rng::Rng::new(12345), whose stream cannot be
    reproduced here (crate not vendored); a SplitMix64-based stream with the same seed is used instead.
    The generated scenes are committed as YAML under scenes/ so every consumer sees identical bytes.
  * examples/part2_all.rs no longer compiles against the reference library (it uses a removed
    ConstantMedium::new signature); it is restated with the current Scene::add_volume API (scene.rs:47-62).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .api import (CameraSettings, CheckerTexture, Cone, ConstantTexture, Cylinder, DielectricMat, Disk,
                  EmissiveMat, F, HdrEnvironment, ImageTexture, LambertianMat, MetalMat, Rect3d, RenderObject,
                  Renderer, Rotor3, Scene, SkyEnv, Sphere, TriangleMesh, TurbulenceTexture, Vec3, XYRect, XZRect, YZRect,
                  to_radians)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE_DIR = os.path.join(REPO, "scenes")


class SceneRng:
    """SplitMix64 -> f32 in [0,1): stands in for tiny_rng::Rng in the example scene generators."""

    def __init__(self, seed: int):
        self.state = seed & 0xFFFFFFFFFFFFFFFF

    def next_u64(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def rand_f32(self):
        return F((self.next_u64() >> 40) * (1.0 / 16777216.0))


# --- examples/random_spheres.rs:14-67 --------------------------------------------------------------
def random_scene(rand: SceneRng) -> Scene:
    scene = Scene.new()
    checker_mat = scene.add_material(LambertianMat(CheckerTexture.with_colors(
        Vec3(0.2, 0.4, 0.1), Vec3(0.9, 0.9, 0.9), 10.0)))
    scene.add_object(RenderObject.new(Sphere(1000.0, checker_mat)).position(0.0, -1000.0, -1.0))
    for x in range(-11, 11):
        for y in range(-11, 11):
            center = Vec3(F(x) + F(0.9) * rand.rand_f32(), 0.2, F(y) + F(0.9) * rand.rand_f32())
            if (center - Vec3(4.0, 0.2, 0.9)).mag() > F(0.9):
                pick = rand.rand_f32()
                if pick < F(0.8):
                    mat = scene.add_material(LambertianMat.with_color(Vec3(
                        rand.rand_f32() * rand.rand_f32(), rand.rand_f32() * rand.rand_f32(),
                        rand.rand_f32() * rand.rand_f32())))
                elif pick < F(0.95):
                    mat = scene.add_material(MetalMat(
                        Vec3(F(0.5) * (F(1.0) + rand.rand_f32()), F(0.5) * (F(1.0) + rand.rand_f32()),
                             F(0.5) * (F(1.0) + rand.rand_f32())), F(0.5) * rand.rand_f32()))
                else:
                    mat = scene.add_material(DielectricMat(1.5))
                scene.add_object(RenderObject.new(Sphere(0.2, mat)).position_vec(center))
    glass = scene.add_material(DielectricMat(1.5))
    diffuse = scene.add_material(LambertianMat.with_color(Vec3(0.4, 0.2, 0.1)))
    metal = scene.add_material(MetalMat(Vec3(0.7, 0.6, 0.5), 0.0))
    scene.add_object(RenderObject.new(Sphere(1.0, glass)).position(0.0, 1.0, 0.0))
    scene.add_object(RenderObject.new(Sphere(1.0, diffuse)).position(-4.0, 1.0, 0.0))
    scene.add_object(RenderObject.new(Sphere(1.0, metal)).position(4.0, 1.0, 0.0))
    scene.set_environment(SkyEnv.default())
    return scene


# --- examples/cornell_box.rs:13-48 -----------------------------------------------------------------
def cornell_box() -> Scene:
    world = Scene.new()
    red = world.add_material(LambertianMat.with_color(Vec3(0.65, 0.05, 0.05)))
    white = world.add_material(LambertianMat.with_color(Vec3(0.73, 0.73, 0.73)))
    green = world.add_material(LambertianMat.with_color(Vec3(0.12, 0.45, 0.15)))
    light = world.add_material(EmissiveMat.with_color(Vec3(15.0, 15.0, 15.0)))
    world.add_object(RenderObject.new(XZRect(213.0, 343.0, 227.0, 332.0, 554.0, light)))
    world.add_object(RenderObject.new(YZRect(0.0, 555.0, 0.0, 555.0, 555.0, green)).flip_normals())
    world.add_object(RenderObject.new(YZRect(0.0, 555.0, 0.0, 555.0, 0.0, red)))
    world.add_object(RenderObject.new(XZRect(0.0, 555.0, 0.0, 555.0, 0.0, white)))
    world.add_object(RenderObject.new(XZRect(0.0, 555.0, 0.0, 555.0, 555.0, white)).flip_normals())
    world.add_object(RenderObject.new(XYRect(0.0, 555.0, 0.0, 555.0, 555.0, white)).flip_normals())
    world.add_object(RenderObject.new(Rect3d.with_size(Vec3(165.0, 165.0, 165.0), white))
                     .rotate(Rotor3.from_rotation_xz(to_radians(18.0))).position(130.0, 0.0, 65.0))
    world.add_object(RenderObject.new(Rect3d.with_size(Vec3(165.0, 330.0, 165.0), white))
                     .rotate(Rotor3.from_rotation_xz(to_radians(-15.0))).position(265.0, 0.0, 295.0))
    return world


# --- examples/earth.rs:12-40 -----------------------------------------------------------------------
def earth_scene() -> Scene:
    scene = Scene.new()
    earth_mat = scene.add_material(LambertianMat(ImageTexture("earthmap.jpg")))
    uv_image_mat = scene.add_material(LambertianMat(ImageTexture("uvmap.png")))
    scene.add_object(RenderObject.new(Sphere(0.25, earth_mat)))
    scene.add_object(RenderObject.new(Sphere(0.25, uv_image_mat)).position(1.0, 0.0, 0.0))
    scene.add_object(RenderObject.new(Sphere(0.25, earth_mat)).position(0.0, 1.0, 0.0))
    scene.add_object(RenderObject.new(Sphere(0.25, earth_mat)).position(0.0, 0.0, 1.0))
    grey = scene.add_material(LambertianMat.with_color(Vec3.broadcast(0.5)))
    scene.add_object(RenderObject.new(XZRect(-100.0, 100.0, -100.0, 100.0, 0.0, grey)))
    light = scene.add_material(EmissiveMat.with_color(Vec3.broadcast(8.0)))
    scene.add_object(RenderObject.new(YZRect(0.0, 20.0, 0.0, 10.0, -3.0, light))
                     .rotate(Rotor3.from_rotation_xz(-30.0)).position(0.0, 0.0, -10.0))
    scene.set_environment(SkyEnv.default())
    return scene


# --- examples/hdri_test.rs:85-103 ------------------------------------------------------------------
HDR_ASSET = "synthetic_hdr:4096x2048:7"  # urban_street_04_4k.hdr is not in the reference repository


def hdri_test() -> Scene:
    scene = Scene.new()
    scene.set_environment(HdrEnvironment.from_path(HDR_ASSET))
    glass = scene.add_material(DielectricMat(1.5))
    diffuse = scene.add_material(LambertianMat.with_color(Vec3(0.8, 0.8, 0.8)))
    metal = scene.add_material(MetalMat(Vec3(0.7, 0.7, 0.7), 0.0))
    scene.add_object(RenderObject.new(Sphere(1.0, glass)).position(0.0, 1.0, 0.0))
    scene.add_object(RenderObject.new(Sphere(1.0, diffuse)).position(-4.0, 1.0, 0.0))
    scene.add_object(RenderObject.new(Sphere(1.0, metal)).position(4.0, 1.0, 0.0))
    scene.add_object(RenderObject.new(XZRect(-100.0, 100.0, -100.0, 100.0, 0.0, diffuse)))
    return scene


# --- examples/volume_test.rs:11-46 -----------------------------------------------------------------
def volume_scene() -> Scene:
    scene = Scene.new()
    glass = scene.add_material(DielectricMat(1.5))
    diffuse = scene.add_material(LambertianMat.with_color(Vec3(0.8, 0.8, 0.8)))
    scene.add_material(MetalMat(Vec3(0.7, 0.7, 0.7), 0.0))
    scene.add_volume(RenderObject.new(Sphere(1.0, diffuse)).position(0.0, 1.0, 0.0), 0.5,
                     ConstantTexture.from_rgb(0.5, 0.0, 0.8))
    scene.add_object(RenderObject.new(Sphere(1.01, glass)).position(0.0, 1.0, 1.0))
    scene.add_object(RenderObject.new(XZRect(-100.0, 100.0, -100.0, 100.0, 0.0, diffuse)))
    light = scene.add_material(EmissiveMat.with_color(Vec3.broadcast(8.0)))
    scene.add_object(RenderObject.new(YZRect(0.0, 20.0, 0.0, 10.0, -3.0, light))
                     .rotate(Rotor3.from_rotation_xz(-30.0)).position(0.0, 0.0, -10.0))
    scene.set_environment(SkyEnv.default())
    return scene


# --- examples/heightmap.rs:11-52 (restated with the current API: Scene::add_mesh no longer exists) ---
def heightmap_scene() -> Scene:
    scene = Scene.new()
    scene.set_environment(SkyEnv.default())
    light = scene.add_material(EmissiveMat.with_color(Vec3.broadcast(8.0)))
    scene.add_object(RenderObject.new(YZRect(0.0, 20.0, 0.0, 10.0, -3.0, light))
                     .rotate(Rotor3.from_rotation_xz(-30.0)).position(0.0, 0.0, -10.0))
    green = scene.add_material(LambertianMat.with_color(Vec3(0.0, 0.5, 0.3)))
    verts, indicies, size = [], [], 20
    for y in range(size):
        for x in range(size):
            height = np.cos(np.float32(x), dtype=np.float32) + np.sin(np.float32(y), dtype=np.float32)
            verts.append((np.float32(x), np.float32(0.2) * np.float32(height), np.float32(y)))
            if x != 0 and y != 0:
                indicies += [(y - 1) * size + x - 1, y * size + x, (y - 1) * size + x]
                indicies += [(y - 1) * size + x - 1, y * size + x - 1, y * size + x]
    scene.add_object(RenderObject.new(TriangleMesh(verts, indicies, None, None, green)))
    return scene


# --- examples/part2_all.rs:13-80 (restated with Scene::add_volume) -------------------------------
def final_scene(rand: SceneRng) -> Scene:
    scene = Scene.new()
    ground = scene.add_material(LambertianMat.with_color(Vec3(0.48, 0.83, 0.53)))
    origin = Vec3(-10.0, 0.0, -10.0)
    for x in range(20):
        for z in range(20):
            pos = origin + Vec3(F(x), 0.0, F(z))
            size = Vec3(1.0, rand.rand_f32() + F(0.01), 1.0)
            scene.add_object(RenderObject.new(Rect3d.with_size(size, ground)).position_vec(pos))
    light = scene.add_material(EmissiveMat.with_color(F(7.0) * Vec3.one()))
    scene.add_object(RenderObject.new(XZRect(1.23, 4.23, 1.47, 4.12, 5.54, light)))
    brown = scene.add_material(LambertianMat.with_color(Vec3(0.7, 0.3, 0.1)))
    scene.add_object(RenderObject.new(Sphere(0.5, brown)).position(4.0, 4.0, 2.0))
    glass = scene.add_material(DielectricMat(1.5))
    scene.add_object(RenderObject.new(Sphere(0.5, glass)).position(2.6, 1.5, 0.45))
    metal = scene.add_material(MetalMat(Vec3(0.8, 0.8, 0.9), 10.0))
    scene.add_object(RenderObject.new(Sphere(0.5, metal)).position(0.0, 1.5, 1.45))
    scene.add_object(RenderObject.new(Sphere(0.7, glass)).position(3.6, 1.5, 1.45))
    scene.add_volume(RenderObject.new(Sphere(0.7, glass)).position(3.6, 1.5, 1.45), 0.2,
                     ConstantTexture(Vec3(0.2, 0.4, 0.9)))
    earth_mat = scene.add_material(LambertianMat(ImageTexture("earthmap.jpg")))
    scene.add_object(RenderObject.new(Sphere(1.0, earth_mat)).position(4.0, 2.0, 4.0))
    noise = scene.add_material(LambertianMat(TurbulenceTexture(5, 10.0)))
    scene.add_object(RenderObject.new(Sphere(0.8, noise)).position(2.2, 2.8, 3.0))
    white = scene.add_material(LambertianMat.with_color(F(0.73) * Vec3.one()))
    for _ in range(1000):
        pos = F(1.65) * Vec3(rand.rand_f32(), rand.rand_f32(), rand.rand_f32()) + Vec3(1.0, 2.7, 3.95)
        scene.add_object(RenderObject.new(Sphere(0.1, white)).position_vec(pos))
    scene.add_volume(RenderObject.new(Sphere(5000.0, 0)), 0.0001, ConstantTexture(Vec3.one()))
    return scene


# ---------------------------------------------------------------------------------------------------
@dataclass
class Config:
    name: str
    scene_file: str                 # under scenes/
    width: int
    height: int
    samples: int
    use_bvh: bool
    cam_pos: tuple
    look_at: tuple
    vfov: float = 30.0
    aperture: float = 0.0
    focus_dist: float = 10.0
    cite: str = ""

    def renderer(self, width=None, height=None, samples=None, seed=0) -> Renderer:
        cam = (CameraSettings.default().cam_pos(self.cam_pos).look_at(self.look_at).field_of_view(self.vfov)
               .aperture(self.aperture).focus_dist(self.focus_dist))
        return (Renderer.default().width(width or self.width).height(height or self.height)
                .samples(samples or self.samples).use_bvh(self.use_bvh).camera(cam).seed(seed))

    def path(self) -> str:
        p = os.path.join(SCENE_DIR, self.scene_file)
        return p if os.path.exists(p) else p + ".gz"

    def load(self) -> Scene:
        s = Scene.from_file(self.path())
        s.asset_dir = os.path.join(SCENE_DIR, "assets")
        return s


# BASELINE.json configs: C1..C5, plus conics / volume as extra parity scenes.
CONFIGS = {
    "random_spheres": Config("random_spheres", "random_spheres.yml", 960, 540, 32, True, (13, 2, 3), (0, 0, 0),
                             vfov=30.0, aperture=0.1, cite="examples/random_spheres.rs:69-91"),
    "cornell_box": Config("cornell_box", "cornell_box.yml", 300, 300, 1024, False, (278, 278, -800), (278, 278, 0),
                          vfov=40.0, cite="examples/cornell_box.rs:50-72 (spp per BASELINE.json)"),
    "suzanne": Config("suzanne", "suzanne.yml", 1920, 1080, 1024, True, (1, 2.5, 5), (0, 0, 0), vfov=40.0,
                      cite="examples/suzanne.rs:86-96 camera; BASELINE.json 1080p x 1024 spp"),
    "teapot": Config("teapot", "teapot.yml", 1920, 1080, 1024, True, (1, 4, 8), (0, 1, 0), vfov=40.0,
                     cite="examples/teapot.rs:99-109 camera; BASELINE.json 1080p x 1024 spp"),
    "hdri_test": Config("hdri_test", "hdri_test.yml", 500, 250, 500, False, (0, 2, -10), (0, 0, 0),
                        cite="examples/hdri_test.rs:105-119"),
    "earth": Config("earth", "earth.yml", 800, 800, 128, False, (5, 5, 5), (0, 0, 0), vfov=30.0,
                    cite="examples/earth.rs:42-55"),
    "part2_all": Config("part2_all", "part2_all.yml", 3840, 2160, 4096, True, (-9, 3, -9), (1, 3, 2), vfov=25.0,
                        cite="examples/part2_all.rs:82-98 camera; BASELINE.json 4K x 4096 spp"),
    "conics": Config("conics", "conics.yml", 960, 540, 128, False, (6, 4, -7), (0, 1.5, 0), vfov=60.0,
                     cite="examples/conics.rs:86-96"),
    "volume": Config("volume", "volume.yml", 960, 540, 2048, False, (0, 2, -10), (0, 0, 0),
                     cite="examples/volume_test.rs:62-70"),
    "heightmap": Config("heightmap", "heightmap.yml", 960, 540, 32, True, (10, 6, -3), (10, 0, 10), vfov=40.0,
                        cite="examples/heightmap.rs:54-66"),
    # the CLI's hard-coded view of any .yml (main.rs:28-38): BVH on, so Disk's degenerate bbox is live
    "conics_cli": Config("conics_cli", "conics.yml", 960, 540, 128, True, (0, 30, 50), (0, 0, 0), vfov=40.0,
                         cite="src/main.rs:28-38"),
}


def generate_scene_files(out_dir: str = SCENE_DIR) -> None:
    """(Re)write the generated scene documents. suzanne/teapot/conics come from the reference's own
    serde dumps (tools/import_reference_scenes.py) and are not regenerated here."""
    os.makedirs(out_dir, exist_ok=True)
    gens = {
        "random_spheres.yml": lambda: random_scene(SceneRng(12345)),
        "cornell_box.yml": cornell_box,
        "earth.yml": earth_scene,
        "hdri_test.yml": hdri_test,
        "volume.yml": volume_scene,
        "heightmap.yml": heightmap_scene,
        "part2_all.yml": lambda: final_scene(SceneRng(12345)),
    }
    for name, fn in gens.items():
        with open(os.path.join(out_dir, name), "w") as f:
            f.write(fn().to_yaml())
