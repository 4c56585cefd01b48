"""CPU tests of the host side of the boundary: the C-ABI library loads and exports every declared symbol,
the C++ YAML loader / BVH builder / flattener agree with the oracle's independent (PyYAML + pointer tree)
construction, the mirrored API serialises the reference's document layout, and compute fails loudly
without a GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ALL_SCENES, CONFIGS, REPO, native_scene, oracle_scene, params_for, scene_doc, scene_text
from firework_b200 import _native as N
from firework_b200.engine import NativeScene
from firework_b200.serde_yaml import dumps, loads


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "firework_b200.h")).read()
    declared = set(re.findall(r"\b(fw_[a-z0-9_]+)\s*\(", header))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    L = N.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.fw_version()


def test_library_has_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out


@pytest.mark.parametrize("name", ALL_SCENES)
def test_bvh_build_matches_oracle(name):
    ns = native_scene(name, commit=False)
    ns.build_host()
    orc = oracle_scene(name, use_bvh=True)
    order, n_nodes, depth = orc.bvh_leaf_order()
    assert np.array_equal(ns.top_leaf_order(), order)                    # bvh.rs:21-71 split + stable sort
    assert np.array_equal(ns.object_aabbs(), orc.object_aabbs())         # scene.rs:167-212 (bit-exact)
    for i, ro in enumerate(scene_doc(name)["render_objects"]):
        if ro["obj"]["object_type"] == "TriangleMesh":
            nt = len(ro["obj"]["indicies"]) // 3
            tri_order, _, _ = orc.mesh_leaf_order(i, nt)
            assert np.array_equal(ns.mesh_leaf_order(i), tri_order)
    ns.close()


def test_yaml_emitter_matches_reference_layout():
    # conics.yml is a serde_yaml dump made by the reference itself: re-emitting its parse must give the same
    # line structure (keys, indentation, sequence style), and identical values
    ref = scene_text("conics")
    doc = loads(ref)
    out = dumps(doc)
    assert loads(out) == doc
    strip = lambda t: [re.sub(r"[-0-9.e]+$", "N", l) for l in t.rstrip().split("\n")]
    assert strip(ref) == strip(out)


def test_mirrored_api_builds_reference_document():
    from firework_b200 import scenes
    doc = loads(scenes.cornell_box().to_yaml())
    assert [o["obj"]["object_type"] for o in doc["render_objects"]] == \
        ["XZRect", "YZRect", "YZRect", "XZRect", "XZRect", "XYRect", "Rect3d", "Rect3d"]
    box = doc["render_objects"][6]
    assert [list(f)[0] for f in box["obj"]["faces"]] == ["XY", "XY", "XZ", "XZ", "YZ", "YZ"]      # rect3d.rs:19-77
    assert [list(f.values())[0]["flip_normal"] for f in box["obj"]["faces"]] == [False, True] * 3
    assert abs(box["rotation"]["bv"]["xz"] + np.sin(np.radians(9.0))) < 1e-6
    vol = loads(scenes.volume_scene().to_yaml())
    med = vol["render_objects"][0]["obj"]
    assert med["object_type"] == "ConstantMedium" and med["obj"]["object_type"] == "Sphere"
    assert vol["materials"][med["material"]]["material"] == "IsotropicMat"                           # scene.rs:47-62


def _load(text):
    h = C.c_void_p()
    data = text.encode()
    return N.lib().fw_scene_from_yaml(data, len(data), C.byref(h)), h


def test_loader_errors_are_reported_not_thrown():
    good = scene_text("cornell_box")
    rc, h = _load(good.replace("object_type: XYRect", "object_type: Torus"))
    assert rc == -1 and b"unknown object_type `Torus`" in N.lib().fw_last_error()
    rc, h = _load(good.replace("material: EmissiveMat", "material: GlowMat"))
    assert rc == -1 and b"unknown material tag" in N.lib().fw_last_error()
    rc, h = _load(good.replace("      k: 554.0\n", "", 1))
    assert rc == -1 and b"missing field `k`" in N.lib().fw_last_error()
    rc, h = _load("render_objects: []\nmaterials: []\nenvironment:\n  environment: ColorEnv\n  color: {x: 0, y: 0, z: 0}\n")
    assert rc == 0
    assert N.lib().fw_scene_build_host(h) == -2 and b"No render objects" in N.lib().fw_last_error()   # scene.rs:161
    N.lib().fw_scene_destroy(h)
    rc, h = _load(good.replace("material: 3", "material: 9", 1))
    assert rc == -1 and b"material 9" in N.lib().fw_last_error()
    rc, h = _load("a: [1, 2\n")
    assert rc == -1


def test_flow_style_and_comments_are_accepted():
    text = """
# hand-written scene
render_objects:
  - obj: {object_type: Sphere, radius: 1.0, material: 0}   # flow mapping
    position: {x: 0.0, y: 1.0, z: 0.0}
    rotation: {s: 1.0, bv: {xy: 0.0, xz: 0.0, yz: 0.0}}
    flip_normals: false
materials:
  - material: LambertianMat
    albedo: {texture: ConstantTexture, color: {x: 0.5, y: 0.5, z: 0.5}}
environment: {environment: ColorEnv, color: {x: 1.0, y: 1.0, z: 1.0}}
"""
    ns = NativeScene(text, commit=False)
    ns.build_host()
    assert ns.num_objects() == 1
    assert np.array_equal(ns.object_aabbs()[0], [-1, 0, -1, 1, 2, 1])


def test_missing_asset_and_missing_gpu_fail_loudly():
    h = C.c_void_p()
    data = scene_text("earth").encode()
    assert N.lib().fw_scene_from_yaml(data, len(data), C.byref(h)) == 0
    assert N.lib().fw_scene_num_assets(h) == 2
    assert N.lib().fw_scene_asset_path(h, 0) == b"earthmap.jpg"
    assert N.lib().fw_scene_commit(h, 0) == -3 and b"earthmap.jpg" in N.lib().fw_last_error()
    N.lib().fw_scene_destroy(h)
    if N.lib().fw_device_count() == 0:
        # no CPU fallback: committing / rendering without a device is an error
        ns = native_scene("cornell_box", commit=False)
        with pytest.raises(N.FireworkError) as e:
            ns.commit()
        assert e.value.code == -4
        from conftest import params_for
        with pytest.raises(N.FireworkError) as e:
            ns.render(params_for("cornell_box", 8, 8, 1))
        assert e.value.code == -6


# ---- asset ingestion: pinned by the reference's own committed artefacts ----------------------
def _yaml_meshes(name):
    import gzip
    from firework_b200.scenes import SCENE_DIR
    from firework_b200.serde_yaml import loads
    doc = loads(gzip.open(os.path.join(SCENE_DIR, f"{name}.yml.gz"), "rt").read())
    return doc, [o["obj"] for o in doc["render_objects"] if o["obj"]["object_type"] == "TriangleMesh"]


@pytest.mark.parametrize("name,with_normals", [("suzanne", False), ("teapot", True)])
def test_obj_loader_reproduces_the_references_scene_dumps(name, with_normals):
    """Loading the reference's suzanne.obj / teapot.obj must give exactly the vertex, normal and index arrays that
    the reference itself serialised into scenes/suzanne.yml / scenes/teapot.yml (examples/suzanne.rs:78-81)."""
    from firework_b200.assets import load_obj
    from firework_b200.scenes import SCENE_DIR
    models = load_obj(os.path.join(SCENE_DIR, "assets", f"{name}.obj"))
    _, meshes = _yaml_meshes(name)
    assert len(models) == len(meshes)
    for m, ref in zip(models, meshes):
        v = np.array([[p["x"], p["y"], p["z"]] for p in ref["verts"]], np.float32)
        assert np.array_equal(m["positions"], v)
        assert np.array_equal(m["indices"], np.array(ref["indicies"], np.uint32))
        if with_normals:
            n = np.array([[p["x"], p["y"], p["z"]] for p in ref["normals"]], np.float32)
            assert np.array_equal(m["normals"], n)
        else:
            assert ref["normals"] is None


def test_add_obj_rebuilds_the_teapot_scene_document():
    """examples/teapot.rs rebuilt through the mirrored API serialises to the same render objects as scenes/teapot.yml."""
    from firework_b200 import api
    from firework_b200.scenes import SCENE_DIR
    doc, _ = _yaml_meshes("teapot")
    sc = api.Scene.new()
    green = sc.add_material(api.LambertianMat(api.ConstantTexture(api.Vec3(0.2, 0.8, 0.3))))
    api.add_obj(sc, os.path.join(SCENE_DIR, "assets", "teapot.obj"), green, with_normals=True, rotate=api.Rotor3.from_rotation_xz(90.0))
    mine = sc.to_dict()["render_objects"]
    ref = [o for o in doc["render_objects"] if o["obj"]["object_type"] == "TriangleMesh"]
    assert len(mine) == len(ref) == 4
    for a, b in zip(mine, ref):
        assert a["obj"]["indicies"] == b["obj"]["indicies"]
        assert a["rotation"] == b["rotation"] and a["position"] == b["position"]
        assert np.array_equal(np.array([[p["x"], p["y"], p["z"]] for p in a["obj"]["verts"]], np.float32),
                              np.array([[p["x"], p["y"], p["z"]] for p in b["obj"]["verts"]], np.float32))


def _write_rgbe(path, rgbe, rle):
    """Minimal Radiance writer for the decoder test: rgbe (H, W, 4) u8; rle = new-style run-length scanlines."""
    h, w, _ = rgbe.shape
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        for y in range(h):
            if not rle:
                f.write(rgbe[y].tobytes())
                continue
            f.write(bytes([2, 2, w >> 8, w & 255]))
            for ch in range(4):
                row, x = rgbe[y, :, ch], 0
                while x < w:
                    run = 1
                    while x + run < w and run < 127 and row[x + run] == row[x]:
                        run += 1
                    if run >= 3:
                        f.write(bytes([128 + run, int(row[x])]))
                        x += run
                    else:
                        n = min(w - x, 100)
                        f.write(bytes([n]) + row[x:x + n].tobytes())
                        x += n


@pytest.mark.parametrize("rle", [False, True])
def test_radiance_hdr_decoder(tmp_path, rle):
    """fw_hdr_load: flat and run-length scanlines, texel -> float as image 0.23.9 (c * 2^(e-136), e == 0 -> black)."""
    from firework_b200.assets import load_hdr
    rng = np.random.default_rng(5)
    h, w = 7, 40
    rgbe = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    rgbe[2, 5:30] = rgbe[2, 5]            # long runs
    rgbe[3, :, 3] = 0                     # e == 0 -> (0, 0, 0)
    rgbe[4, :, 3] = rng.integers(100, 150, size=w)
    p = str(tmp_path / "t.hdr")
    _write_rgbe(p, rgbe, rle)
    got = load_hdr(p)
    e = rgbe[..., 3].astype(np.float32)
    want = np.where(e[..., None] == 0, np.float32(0), np.exp2(e - np.float32(136))[..., None] * rgbe[..., :3].astype(np.float32)).astype(np.float32)
    assert got.shape == (h, w, 3) and np.array_equal(got, want)


def test_asset_loaders_report_errors(tmp_path):
    from firework_b200._native import FireworkError
    from firework_b200.assets import load_hdr, load_obj
    with pytest.raises(N.FireworkError):
        load_obj(str(tmp_path / "missing.obj"))
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")
    with pytest.raises(N.FireworkError, match="bad vertex index"):
        load_obj(str(bad))
    nothdr = tmp_path / "x.hdr"
    nothdr.write_bytes(b"P6\n1 1\n255\n\0\0\0")
    with pytest.raises(N.FireworkError, match="not a Radiance"):
        load_hdr(str(nothdr))


def _top_tree_leaves(ns):
    """Walk the flattened top-level tree: {leaf code: (box lo, box hi)} and the boxes of every interior slot with the
    leaf boxes below it (for the containment check)."""
    nodes, root = ns.bvh_nodes()
    codes = nodes[:, 6, :].view(np.int32)
    leaves, checks = {}, []

    def walk(code):
        """returns the list of (lo, hi) leaf boxes below `code`"""
        below = []
        for k in range(4):
            c = int(codes[code, k])
            if c == -2147483648:
                continue  # empty slot
            lo, hi = nodes[code, 0:3, k].copy(), nodes[code, 3:6, k].copy()
            if c >= 0:
                sub = walk(c)
                checks.append((lo, hi, sub))
                below += sub
            else:
                assert c not in leaves
                leaves[c] = (lo, hi)
                below.append((lo, hi))
        return below

    if root >= 0:
        walk(root)
    return leaves, checks


@pytest.mark.parametrize("name", ["random_spheres", "part2_all", "teapot"])
def test_reclustered_bvh_keeps_the_references_leaves(name):
    """The SAH interior must sit on exactly the reference's leaves: same leaf codes (item ranges in DFS order), same
    leaf boxes bit for bit as the tree built with the reference's own interior (FW_BVH_SAH=0, separate process), every
    item covered once, and every interior slot box must contain all the leaf boxes below it."""
    import subprocess
    import sys
    ns = native_scene(name, commit=False)
    ns.build_host()
    leaves, checks = _top_tree_leaves(ns)
    n_obj = ns.num_objects()
    ns.close()
    covered = []
    for c in leaves:
        p = ~c
        covered += list(range(p >> 1, (p >> 1) + (p & 1) + 1))
    assert sorted(covered) == list(range(n_obj))
    for lo, hi, sub in checks:
        for llo, lhi in sub:
            assert np.all(lo <= llo) and np.all(lhi <= hi)
    prog = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from test_host import _top_tree_leaves; from conftest import native_scene\n"
            "ns = native_scene(%r, commit=False); ns.build_host(); l, _ = _top_tree_leaves(ns)\n"
            "print(json.dumps({str(k): [v[0].tolist(), v[1].tolist()] for k, v in l.items()}))\n"
            % (REPO, os.path.join(REPO, "tests"), name))
    out = subprocess.run([sys.executable, "-c", prog], env=dict(os.environ, FW_BVH_SAH="0"), capture_output=True, text=True, check=True).stdout
    import json
    ref = json.loads(out.strip().splitlines()[-1])
    assert set(ref) == {str(k) for k in leaves}
    for k, (lo, hi) in leaves.items():
        assert np.array_equal(np.array(ref[str(k)][0], np.float32), lo) and np.array_equal(np.array(ref[str(k)][1], np.float32), hi)


def test_hostile_inputs_are_rejected_not_crashed(tmp_path):
    """Untrusted inputs must come back as error codes: deeply nested flow YAML (recursion), a Radiance header claiming an
    absurd size (allocation overflow), a render call on a scene that was never committed (no context to read)."""
    import ctypes as C
    from firework_b200._native import FireworkError, FwParams, FwStats
    from firework_b200.assets import load_hdr
    from firework_b200.engine import NativeScene
    with pytest.raises(N.FireworkError, match="nested too deeply"):
        NativeScene("render_objects: " + "[" * 100000 + "]" * 100000 + "\nmaterials: []\n", commit=False)
    big = tmp_path / "big.hdr"
    big.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2147483648 +X 2147483648\n" + b"\0" * 64)
    with pytest.raises(N.FireworkError, match="unreasonable image size"):
        load_hdr(str(big))
    from conftest import params_for
    ns = native_scene("cornell_box", commit=False)
    p = params_for("cornell_box", 8, 8, 1)
    st = FwStats()
    rc = N.lib().fw_render_accumulate_device(ns._h, C.byref(p), C.c_void_p(16), None, C.byref(st))
    assert rc == -6 and b"commit" in N.lib().fw_last_error()          # FW_ERR_STATE, not a crash
    rc = N.lib().fw_render_multi(ns._h, C.byref(p), 2, None, 0, None, None, C.byref(st), None)
    assert rc == -6
    ns.close()


def test_static_library_is_built_beside_the_shared_one():
    """BASELINE.json north_star: "a CUDA static library built by build.rs" — the archive a Rust host would link."""
    import subprocess
    from firework_b200.build import STATIC_LIB
    assert os.path.exists(STATIC_LIB)
    members = subprocess.run(["ar", "t", STATIC_LIB], capture_output=True, text=True, check=True).stdout.split()
    assert {"api.o", "kernels_extend.o", "kernels_shade.o", "scene_host.o", "multi_gpu.o"} <= set(members)
    syms = subprocess.run(["nm", "-g", "--defined-only", STATIC_LIB], capture_output=True, text=True, check=True).stdout
    for name in ("fw_scene_from_yaml", "fw_render", "fw_render_multi"):
        assert re.search(rf"\bT {name}\b", syms), name


def test_checkpoint_fingerprint_guards_resume(tmp_path):
    """A checkpoint may only be resumed by the render that wrote it (same scene / camera / renderer parameters)."""
    from firework_b200.api import Scene
    from firework_b200.progressive import render_fingerprint
    a = Scene.from_file(CONFIGS["cornell_box"].path())
    b = Scene.from_file(CONFIGS["volume"].path())
    ra = CONFIGS["cornell_box"].renderer(width=16, height=16, samples=4, seed=1)
    assert render_fingerprint(a, ra) == render_fingerprint(a, CONFIGS["cornell_box"].renderer(width=16, height=16, samples=9, seed=1))
    assert render_fingerprint(a, ra) != render_fingerprint(b, ra)
    assert render_fingerprint(a, ra) != render_fingerprint(a, CONFIGS["cornell_box"].renderer(width=16, height=16, samples=4, seed=2))
    assert render_fingerprint(a, ra) != render_fingerprint(a, CONFIGS["cornell_box"].renderer(width=17, height=16, samples=4, seed=1))


# ---- image files: PNG / JPEG in (ImageTexture::from_path, texture.rs:285-292), PNG out (window.rs save_image) -------------------
def test_native_png_decoder_matches_pil_on_every_colour_type(tmp_path):
    from PIL import Image
    from firework_b200.assets import load_image_native
    from firework_b200.scenes import SCENE_DIR
    p = os.path.join(SCENE_DIR, "assets", "uvmap.png")
    assert np.array_equal(load_image_native(p), np.asarray(Image.open(p).convert("RGBA")))
    rng = np.random.default_rng(3)
    smooth = (np.add.outer(np.arange(37), np.arange(53)) * 3 % 256).astype(np.uint8)      # exercises the Sub / Up / Avg / Paeth filters
    cases = {
        "rgb": Image.fromarray(np.stack([smooth, smooth.T[:37, :53] if False else smooth[::-1], rng.integers(0, 256, smooth.shape, dtype=np.uint8)], -1), "RGB"),
        "rgba": Image.fromarray(rng.integers(0, 256, (19, 23, 4), dtype=np.uint8), "RGBA"),
        "gray": Image.fromarray(smooth, "L"),
        "gray_alpha": Image.fromarray(rng.integers(0, 256, (11, 7, 2), dtype=np.uint8), "LA"),
        "palette": Image.fromarray(rng.integers(0, 256, (31, 17, 3), dtype=np.uint8), "RGB").quantize(16),
        "bilevel": Image.fromarray((rng.random((13, 29)) > 0.5)),
    }
    for name, im in cases.items():
        f = str(tmp_path / f"{name}.png")
        im.save(f, optimize=(name != "rgb"))
        got = load_image_native(f)
        want = np.asarray(Image.open(f).convert("RGBA"))
        assert got.shape == want.shape, name
        assert np.array_equal(got, want), name


def test_native_jpeg_decoder_is_within_rounding_of_libjpeg():
    """A JPEG decoder is defined up to its IDCT rounding and chroma upsampling; this one (accurate IDCT, triangle upsampling,
    libjpeg's fixed-point colour conversion) agrees with PIL's libjpeg on 98 % of the values of earthmap.jpg, max |diff| 3."""
    from PIL import Image
    from firework_b200.assets import load_image_native
    from firework_b200.scenes import SCENE_DIR
    p = os.path.join(SCENE_DIR, "assets", "earthmap.jpg")
    got = load_image_native(p).astype(int)
    want = np.asarray(Image.open(p).convert("RGBA")).astype(int)
    assert got.shape == want.shape
    d = np.abs(got - want)
    assert d.max() <= 4 and (d > 0).mean() < 0.03 and d.mean() < 0.03


def test_native_jpeg_decoder_sampling_modes_and_restarts(tmp_path):
    from PIL import Image
    from firework_b200.assets import load_image_native
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:61, 0:83]
    img = np.stack([(xx * 3) % 256, (yy * 4) % 256, ((xx + yy) * 2) % 256], -1).astype(np.uint8)
    img = (img * 0.8 + rng.integers(0, 50, img.shape)).astype(np.uint8)
    for sub, name in ((0, "444"), (1, "422"), (2, "420")):
        f = str(tmp_path / f"s{name}.jpg")
        Image.fromarray(img, "RGB").save(f, quality=90, subsampling=sub)
        got = load_image_native(f).astype(int)
        want = np.asarray(Image.open(f).convert("RGBA")).astype(int)
        d = np.abs(got - want)
        assert got.shape == want.shape and d.max() <= 4 and d.mean() < 0.2, (name, d.max(), d.mean())
    f = str(tmp_path / "restart.jpg")
    try:
        Image.fromarray(img, "RGB").save(f, quality=90, subsampling=2, restart_marker_blocks=3)
        has_restart = b"\xff\xdd" in open(f, "rb").read()
    except TypeError:
        has_restart = False
    if has_restart:   # DRI + RSTn markers (Pillow >= 10.2 writes them on request)
        d = np.abs(load_image_native(f).astype(int) - np.asarray(Image.open(f).convert("RGBA")).astype(int))
        assert d.max() <= 4 and d.mean() < 0.2
    f = str(tmp_path / "gray.jpg")
    Image.fromarray(img[..., 0], "L").save(f, quality=85)
    d = np.abs(load_image_native(f).astype(int) - np.asarray(Image.open(f).convert("RGBA")).astype(int))
    assert d.max() <= 2
    f = str(tmp_path / "prog.jpg")
    Image.fromarray(img, "RGB").save(f, progressive=True)
    with pytest.raises(N.FireworkError, match="progressive"):
        load_image_native(f)


def test_png_writer_round_trips(tmp_path):
    from PIL import Image
    from firework_b200.assets import load_image_native, write_png
    rng = np.random.default_rng(9)
    rgb = rng.integers(0, 256, (45, 67, 3), dtype=np.uint8)
    f = str(tmp_path / "out.png")
    write_png(f, rgb)
    assert np.array_equal(np.asarray(Image.open(f).convert("RGB")), rgb)
    assert np.array_equal(load_image_native(f)[..., :3], rgb)
    with pytest.raises(N.FireworkError):
        load_image_native(str(tmp_path / "missing.png"))
    open(tmp_path / "junk.png", "wb").write(b"\x89PNG\r\n\x1a\n" + b"\0" * 40)
    with pytest.raises(N.FireworkError):
        load_image_native(str(tmp_path / "junk.png"))


def test_native_cli_mirrors_main_rs_arguments(tmp_path):
    """csrc/cli_main.cpp -> firework_b200/bin/firework: src/main.rs:6-20 options; errors are messages + exit code 1, and without
    a CUDA device the render refuses to start (no CPU fallback)."""
    import subprocess
    from firework_b200.build import CLI
    assert os.path.exists(CLI)
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--scene-file" in r.stdout and "--samples" in r.stdout
    r = subprocess.run([CLI, "-s", "4"], capture_output=True, text=True)
    assert r.returncode == 1 and "--scene-file" in r.stderr
    r = subprocess.run([CLI, "--scene-file", str(tmp_path / "nope.yml"), "-s", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open scene file" in r.stderr
    bad = tmp_path / "bad.yml"
    bad.write_text("render_objects: 7\n")
    r = subprocess.run([CLI, "--scene-file", str(bad), "-s", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("error: scene")
    if N.lib().fw_device_count() == 0:
        r = subprocess.run([CLI, "--scene-file", CONFIGS["conics"].path(), "-s", "1", "-o", str(tmp_path / "x.png")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stderr and not os.path.exists(tmp_path / "x.png")


def _build_example(name, out_dir):
    """g++ against the static archive: the link line of INTEGRATION.md."""
    import subprocess
    from firework_b200.build import STATIC_LIB
    exe = os.path.join(str(out_dir), name)
    cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-ffp-contract=off", "-o", exe, os.path.join(REPO, "examples", name + ".cpp"), STATIC_LIB,
                        "-L" + cuda_lib, "-lcudart_static", "-lz", "-ldl", "-lpthread", "-lrt"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_api_mirror_serialises_the_examples_like_the_python_mirror(tmp_path):
    """include/firework.hpp (C++ mirror of the crate's builder API) and examples/*.cpp (the crate's examples restated with it):
    `Scene::to_yaml` must be the document the Python mirror writes for the same example — byte for byte."""
    import subprocess
    from firework_b200 import scenes
    for name, make in (("cornell_box", scenes.cornell_box), ("earth", scenes.earth_scene), ("volume_test", scenes.volume_scene),
                       ("random_spheres", lambda: scenes.random_scene(scenes.SceneRng(12345))),
                       ("part2_all", lambda: scenes.final_scene(scenes.SceneRng(12345)))):
        exe = _build_example(name, tmp_path)
        r = subprocess.run([exe, "--yaml"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout == make().to_yaml(), name
        if name in ("random_spheres", "part2_all"):     # ... which is the committed scene file the benchmarks load
            assert r.stdout == open(CONFIGS[name].path()).read(), name
        h = C.c_void_p()
        data = r.stdout.encode()
        assert N.lib().fw_scene_from_yaml(data, len(data), C.byref(h)) == 0      # and the native loader takes it
        N.lib().fw_scene_destroy(h)


def test_native_image_decoders_survive_damaged_files(tmp_path):
    """Truncated and bit-flipped PNG / JPEG files must come back as an error (or as some image), never as a crash or a hang:
    the decoders sit on the path of untrusted scene assets."""
    from firework_b200.assets import load_image_native
    from firework_b200.scenes import SCENE_DIR
    rng = np.random.default_rng(11)
    for name in ("uvmap.png", "earthmap.jpg"):
        data = open(os.path.join(SCENE_DIR, "assets", name), "rb").read()
        f = str(tmp_path / ("damaged_" + name))
        outcomes = {"error": 0, "image": 0}
        cases = [data[:n] for n in (0, 1, 7, 8, 20, 33, 100, 1000, len(data) // 2, len(data) - 1)]
        for _ in range(60):
            b = bytearray(data)
            for _k in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(0, min(len(b), 4096) if rng.random() < 0.7 else len(b)))   # headers and tables are up front
                b[pos] = int(rng.integers(0, 256))
            cases.append(bytes(b))
        for c in cases:
            open(f, "wb").write(c)
            try:
                img = load_image_native(f)
                assert img.ndim == 3 and img.shape[2] == 4 and img.shape[0] * img.shape[1] <= 1 << 26
                outcomes["image"] += 1
            except N.FireworkError:
                outcomes["error"] += 1
        assert outcomes["error"] >= 10, (name, outcomes)      # every truncation is an error


def test_native_checkpoint_format_round_trips(tmp_path):
    from firework_b200.progressive import fwck_fingerprint, fwck_load, fwck_save
    rng = np.random.default_rng(2)
    sums = rng.random((9, 14, 3), dtype=np.float32)
    f = str(tmp_path / "x.fwck")
    fwck_save(f, sums, 37, 5, 0x1234567890ABCDEF)
    got, done, seed, fp = fwck_load(f)
    assert np.array_equal(got, sums) and (done, seed, fp) == (37, 5, 0x1234567890ABCDEF)
    assert os.path.getsize(f) == 40 + sums.nbytes
    p = params_for("conics", 240, 135, 10, seed=7)
    a = fwck_fingerprint("scene text", p)
    assert a == fwck_fingerprint("scene text", params_for("conics", 240, 135, 99, seed=7))        # the sample count is not part of it
    assert a != fwck_fingerprint("scene text", params_for("conics", 240, 135, 10, seed=8))
    assert a != fwck_fingerprint("scene text.", p)
    open(f, "wb").write(b"FWCKPT01" + b"\0" * 10)
    with pytest.raises(ValueError):
        fwck_load(f)


def test_cpp_mesh_examples_reproduce_the_references_own_scene_dumps(tmp_path):
    """examples/suzanne.rs and examples/teapot.rs write `serde_yaml::to_string(&scene)` to scenes/suzanne.yml / teapot.yml — the
    files the reference commits.  The same examples restated in C++ (include/firework.hpp + the `add_obj` helper over
    fw_obj_load) must produce a document that parses to exactly that: every vertex, normal, index, rotor component and material
    (the two serialisers only differ in how many digits they print for an f32)."""
    import subprocess
    from firework_b200.scenes import SCENE_DIR
    for name in ("suzanne", "teapot"):
        exe = _build_example(name, tmp_path)
        r = subprocess.run([exe, "--obj", os.path.join(SCENE_DIR, "assets", name + ".obj"), "--yaml"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert loads(r.stdout) == scene_doc(name), name
    r = subprocess.run([exe, "--obj", str(tmp_path / "missing.obj"), "--yaml"], capture_output=True, text=True)
    assert r.returncode == 1 and "add_obj" in r.stderr
