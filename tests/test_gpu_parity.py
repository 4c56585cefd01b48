"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Gates (BASELINE.json north_star):
  1. first-hit object / primitive id per ray bit-exact; t and normal within 1e-5 relative;
  2. one scatter step given identical uniforms within 1e-5;
  3. converged image (4096 spp) >= 40 dB PSNR and mean relative error < 1 %.
Plus stage-level checks (primary rays, textures, environment, low-spp renders, determinism, batch-split and
sample-shard invariance) and size-independent properties at BASELINE's full sizes.
"""
import numpy as np
import pytest

from conftest import ALL_SCENES, native_scene, oracle_scene, params_for, psnr_u8, scene_doc
from firework_b200.scenes import CONFIGS

pytestmark = pytest.mark.gpu

REL = 1e-5  # the tolerance BASELINE.json states for t / normal / scatter
# agreement with the reference's own committed renders (8x8 box means; measured: see the test's print)
#                suzanne  teapot  cornell_box  conics  earth
# measured dB      51.0    45.2      44.2      49.5    43.3
# mean abs /255    0.49    0.87      1.15      0.51    1.37
PNG_PSNR_MIN = {"suzanne": 45.0, "teapot": 40.0, "cornell_box": 40.0, "conics": 44.0, "earth": 40.0, "heightmap": 43.0}
PNG_MEAN_ABS_MAX = {"suzanne": 1.5, "teapot": 2.0, "cornell_box": 2.5, "conics": 1.5, "earth": 2.5, "heightmap": 1.5}


def _rel(a, b, floor=1e-20):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


@pytest.fixture(scope="module")
def scenes():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = (native_scene(name), oracle_scene(name, fast=True))
        return cache[name]

    yield get
    for ns, _ in cache.values():
        ns.close()


def test_shared_division_is_ieee():
    """The linear-scan kernels divide by a ray's direction components through a shared reciprocal
    (intersect.cuh div_by): in its fast domain the quotient must be the IEEE one bit for bit, 2^32 operand pairs."""
    from firework_b200.engine import selftest_shared_division
    for seed in (1, 2):
        bad, bad_tiny = selftest_shared_division(1 << 31, seed)
        assert bad == 0 and bad_tiny == 0, (seed, bad, bad_tiny)


@pytest.mark.parametrize("name", ALL_SCENES)
def test_primary_rays_bit_exact(scenes, name):
    ns, orc = scenes(name)
    p = params_for(name, 160, 90, 4, seed=11)
    for s in (0, 3):
        go, gd = ns.primary_rays(p, s)
        oo, od = orc.primary_rays(p, s)
        assert np.array_equal(go, oo) and np.array_equal(gd, od)


def _ray_generations(orc, p, generations=3):
    """Primary rays, then rays leaving the oracle's hit points along normal + a fixed pseudo-random offset."""
    o, d = orc.primary_rays(p, 0)
    for gen in range(generations):
        yield gen, o, d
        h = orc.first_hit(o, d, seed=p.seed, pixel=np.arange(len(o), dtype=np.uint32), bounce=np.full(len(o), gen, np.uint32))
        hit = h["obj"] >= 0
        rng = np.random.default_rng(gen)
        off = rng.uniform(-0.7, 0.7, size=o.shape).astype(np.float32)
        nrm = h["normal"] / np.maximum(np.linalg.norm(h["normal"], axis=1, keepdims=True), 1e-20)
        o = np.where(hit[:, None], h["point"], o).astype(np.float32)
        d = np.where(hit[:, None], nrm + off, d).astype(np.float32)


@pytest.mark.parametrize("name", ALL_SCENES)
def test_first_hit_gate(scenes, name):
    """Gate 1: ids bit-exact, t / normal / point within 1e-5 relative, on primary and two secondary generations."""
    ns, orc = scenes(name)
    cfg = CONFIGS[name]
    p = params_for(name, 320, 180, 1, seed=1234)
    for gen, o, d in _ray_generations(orc, p):
        n = len(o)
        keys = dict(pixel=np.arange(n, dtype=np.uint32), sample=np.zeros(n, np.uint32), bounce=np.full(n, gen, np.uint32))
        g = ns.first_hit(o, d, cfg.use_bvh, seed=1234, **keys)
        r = orc.first_hit(o, d, seed=1234, **keys)
        assert np.array_equal(g["obj"], r["obj"]), f"gen {gen}: {(g['obj'] != r['obj']).sum()} object ids differ"
        assert np.array_equal(g["prim"], r["prim"])
        assert np.array_equal(g["material"], r["material"])
        hit = r["obj"] >= 0
        assert hit.any()
        assert _rel(g["t"][hit], r["t"][hit]).max() <= REL
        nscale = np.maximum(np.linalg.norm(r["normal"][hit], axis=1, keepdims=True), 1e-20)
        nn = np.abs(g["normal"][hit] - r["normal"][hit]) / nscale
        assert np.nanmax(nn) <= REL
        pscale = np.maximum(np.linalg.norm(r["point"][hit], axis=1, keepdims=True), 1.0)
        assert (np.abs(g["point"][hit] - r["point"][hit]) / pscale).max() <= REL
        # uv goes through atan2/asin/acos: a few ulp of [0,1]; NaN uv (asin of 1+eps, sphere.rs:24) must agree too
        assert np.array_equal(np.isnan(g["uv"]), np.isnan(r["uv"]))
        assert np.nanmax(np.abs(g["uv"][hit] - r["uv"][hit]), initial=0.0) <= 1e-5
        # the ordered traversal does no more work than the reference's exhaustive one
        assert g["node_tests"] <= r["aabb_tests"] and g["prim_tests"] <= r["prim_tests"]
        # the same rays through the kernels a render launches (segmented queues, the scene's extend plan — lock-step BVH,
        # linear program or pass 1 + mesh walk + classification —, finalize_hit): identical to the probe kernel bit for bit,
        # hence to the oracle within the same tolerances
        w = ns.first_hit_wavefront(o, d, cfg.use_bvh, seed=1234, sample=0, bounce=gen)
        assert np.array_equal(w["obj"], r["obj"]) and np.array_equal(w["prim"], r["prim"]) and np.array_equal(w["material"], r["material"])
        assert _rel(w["t"][hit], r["t"][hit]).max() <= REL
        assert np.nanmax(np.abs(w["normal"][hit] - r["normal"][hit]) / nscale) <= REL
        assert (np.abs(w["point"][hit] - r["point"][hit]) / pscale).max() <= REL
        assert np.array_equal(np.isnan(w["uv"]), np.isnan(r["uv"]))
        assert np.nanmax(np.abs(w["uv"][hit] - r["uv"][hit]), initial=0.0) <= 1e-5


def test_first_hit_ties_and_coincident_geometry(scenes):
    """bvh.rs:128,141: equal t -> the later leaf wins.  Two coincident spheres / rects in one scene, under BVH
    and linear scan, must resolve to the oracle's object id."""
    from firework_b200.api import (LambertianMat, RenderObject, Scene, Sphere, Vec3, XYRect)
    from firework_b200.engine import NativeScene
    from oracle.oracle import OracleScene
    s = Scene.new()
    m = s.add_material(LambertianMat.with_color(Vec3(0.5, 0.5, 0.5)))
    for k in range(3):   # three identical spheres and three identical rects, interleaved with decoys
        s.add_object(RenderObject.new(Sphere(1.0, m)).position(0.0, 0.0, 0.0))
        s.add_object(RenderObject.new(XYRect(-3.0, 3.0, -3.0, 3.0, -2.0, m)))
        s.add_object(RenderObject.new(Sphere(0.3, m)).position(4.0 + k, 0.0, 0.0))
    rng = np.random.default_rng(5)
    o = np.tile(np.array([[0.0, 0.0, 6.0]], np.float32), (4096, 1))
    d = np.concatenate([rng.uniform(-0.5, 0.5, (4096, 2)), -np.ones((4096, 1))], 1).astype(np.float32)
    ns = NativeScene(s.to_yaml())
    for use_bvh in (True, False):
        orc = OracleScene(s.to_dict(), use_bvh)
        g = ns.first_hit(o, d, use_bvh)
        r = orc.first_hit(o, d)
        assert np.array_equal(g["obj"], r["obj"]) and np.array_equal(g["t"], r["t"])
        assert len(set(r["obj"][r["obj"] >= 0].tolist())) >= 2
    ns.close()


@pytest.mark.parametrize("name", ["teapot", "suzanne", "cornell_box", "conics", "conics_cli", "part2_all",
                                  "random_spheres", "volume"])
def test_nan_direction_rays_follow_the_reference(scenes, name):
    """A ray whose direction is NaN (it happens: scattering off a zero interpolated normal) passes every slab test
    and is accepted by every comparison-rejecting primitive; the reference's merge rules then make the LAST item in
    DFS / scene order win.  The CUDA path answers from a precomputed table instead of walking the whole tree; the
    oracle simply executes the reference logic."""
    ns, orc = scenes(name)
    cfg = CONFIGS[name]
    n = 64
    rng = np.random.default_rng(9)
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    o[n // 2:] = np.nan                       # later bounces: origin is NaN as well
    d = np.full((n, 3), np.nan, np.float32)
    g = ns.first_hit(o, d, cfg.use_bvh)
    r = orc.first_hit(o, d)
    assert np.array_equal(g["obj"], r["obj"]) and np.array_equal(g["prim"], r["prim"])
    assert np.array_equal(g["material"], r["material"])
    hit = r["obj"] >= 0
    assert np.isnan(g["t"][hit]).all() and np.isnan(r["t"][hit]).all()
    assert np.array_equal(np.isnan(g["normal"]), np.isnan(r["normal"]))


@pytest.mark.parametrize("name", ["random_spheres", "cornell_box", "earth", "part2_all", "volume", "hdri_test"])
def test_scatter_step_gate(scenes, name):
    """Gate 2: emit + one scatter step of every material in the scene, identical explicit uniforms, within 1e-5."""
    ns, orc = scenes(name)
    doc = scene_doc(name)
    nmat = len(doc["materials"])
    n = 4096
    rng = np.random.default_rng(17)
    mats = (np.arange(n) % nmat).astype(np.int32)
    ray_d = rng.normal(size=(n, 3)).astype(np.float32) * rng.uniform(0.2, 10.0, (n, 1)).astype(np.float32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    flip = (np.sum(nrm * ray_d, axis=1) > 0) & (rng.uniform(size=n) < 0.7)   # mostly front-facing, some from inside
    nrm[flip] *= -1
    point = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    uv = rng.uniform(0, 1, (n, 2)).astype(np.float32)
    hit_t = rng.uniform(0.1, 10, n).astype(np.float32)
    uniforms = rng.uniform(0, 1, (n, 96)).astype(np.float32)
    uniforms = (np.floor(uniforms * 16777216.0) / 16777216.0).astype(np.float32)   # values the RNG can produce
    g = ns.scatter_step(mats, point - ray_d, ray_d, hit_t, point, nrm, uv, uniforms)
    r = orc.scatter_step(mats, point - ray_d, ray_d, hit_t, point, nrm, uv, uniforms)
    assert (r["consumed"] >= 0).all()
    # the dielectric coin flip compares u with schlick(powf): allow the (rare) knife-edge flips, none expected
    same = g["scattered"] == r["scattered"]
    assert same.all()
    assert np.array_equal(g["consumed"], r["consumed"])
    scale = lambda x: np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-3)
    assert (np.abs(g["d"] - r["d"]) / scale(r["d"])).max() <= REL
    assert (np.abs(g["o"] - r["o"]) / scale(r["o"])).max() <= REL
    # albedo may go through sin / Perlin / image lookups
    assert (np.abs(g["atten"] - r["atten"]) / np.maximum(np.abs(r["atten"]), 1.0)).max() <= REL
    assert (np.abs(g["emit"] - r["emit"]) / np.maximum(np.abs(r["emit"]), 1.0)).max() <= REL


def test_texture_and_environment_lookups(scenes):
    rng = np.random.default_rng(3)
    n = 20000
    pts = rng.uniform(-12, 12, (n, 3)).astype(np.float32)
    uvs = rng.uniform(-0.1, 1.1, (n, 2)).astype(np.float32)   # includes out-of-range uv (clamped: texture.rs:301-302)
    for name, mat in [("random_spheres", 0), ("earth", 0), ("earth", 1), ("part2_all", 7), ("part2_all", 6)]:
        ns, orc = scenes(name)
        tex = ns.material_texture(mat)
        assert tex == orc.tex_of_material[mat] and tex >= 0
        g, r = ns.texture_sample(tex, uvs, pts), orc.texture_sample(tex, uvs, pts)
        bad = np.abs(g - r).max(axis=1) > 1e-5
        # checker flips sign where sin() products straddle zero within an ulp; image / perlin are exact
        assert bad.mean() <= (2e-3 if name == "random_spheres" else 0.0), (name, mat, bad.sum())
    dirs = rng.normal(size=(n, 3)).astype(np.float32)
    for name in ("hdri_test", "random_spheres", "cornell_box"):
        ns, orc = scenes(name)
        g, r = ns.env_sample(dirs), orc.env_sample(dirs)
        bad = np.abs(g - r).max(axis=1) > 1e-5 * np.maximum(np.abs(r).max(axis=1), 1.0)
        assert bad.mean() <= (1e-3 if name == "hdri_test" else 0.0)   # texel-boundary flips from atan2/asin ulps


@pytest.mark.parametrize("name", ALL_SCENES)
def test_low_spp_render_matches_oracle(scenes, name):
    """Same counter-based RNG on both sides => the renders agree path by path; the only divergences come from
    libm-vs-CUDA transcendentals flipping a discrete decision (checker sign, texel index, Schlick coin)."""
    ns, orc = scenes(name)
    p = params_for(name, 160, 90, 4, seed=99)
    grgb, gsum, gst = ns.render(p)
    orgb, osum, ost = orc.render(p)
    assert gst["samples"] == 160 * 90 * 4
    close = np.all(np.abs(gsum - osum) <= 1e-4 * np.maximum(np.abs(osum), 1e-3), axis=2)
    assert close.mean() >= 0.995, f"{(~close).sum()} of {close.size} pixels differ"
    assert abs(int(gst["rays"]) - int(ost["rays"])) <= 1e-3 * ost["rays"]
    d8 = np.abs(grgb.astype(int) - orgb.astype(int)).max(axis=2)
    assert (d8 <= 1).mean() >= 0.995
    assert abs(gsum.mean() - osum.mean()) <= 2e-3 * abs(osum.mean())


def test_render_is_deterministic_and_batch_invariant(scenes):
    ns, _ = scenes("random_spheres")
    p = params_for("random_spheres", 200, 120, 6, seed=4)
    _, a, _ = ns.render(p)
    _, b, _ = ns.render(p)
    assert np.array_equal(a, b)
    for paths in (1000, 24000, 50000):      # pixel tiles, sample chunks, ragged tails
        ns.set_batch_paths(paths)
        _, c, _ = ns.render(p)
        assert np.array_equal(a, c), paths
    ns.set_batch_paths(0)


def test_sample_range_sharding_properties(scenes):
    ns, _ = scenes("cornell_box")
    full = params_for("cornell_box", 128, 128, 8, seed=21)
    _, whole, st = ns.render(full)
    parts = []
    for begin, count in ((0, 3), (3, 5)):
        _, s, _ = ns.render(params_for("cornell_box", 128, 128, 8, seed=21, sample_begin=begin, sample_count=count))
        parts.append(s)
    assert np.allclose(parts[0] + parts[1], whole, rtol=1e-5, atol=1e-6)
    # two single-sample shards add up bit-exactly to the 2-spp render: (0+s0)+s1 == s0+s1
    _, two, _ = ns.render(params_for("cornell_box", 128, 128, 2, seed=21))
    _, s0, _ = ns.render(params_for("cornell_box", 128, 128, 2, seed=21, sample_begin=0, sample_count=1))
    _, s1, _ = ns.render(params_for("cornell_box", 128, 128, 2, seed=21, sample_begin=1, sample_count=1))
    assert np.array_equal(s0 + s1, two)
    # a zero-sample shard contributes nothing
    _, z, _ = ns.render(params_for("cornell_box", 128, 128, 8, seed=21, sample_begin=8, sample_count=0))
    assert not z.any()


def test_device_accumulate_and_resolve_via_torch(scenes):
    """The device-resident entry points used by the multi-GPU path (fw_render_accumulate_device /
    fw_resolve_device) agree with fw_render."""
    import torch
    from firework_b200.distributed import GpuShardRenderer, render_sharded
    ns, _ = scenes("cornell_box")
    r = CONFIGS["cornell_box"].renderer(width=96, height=96, samples=6, seed=8)
    rgb, s, _ = ns.render(r.params())
    gs = GpuShardRenderer(ns, r, 0)
    # emulate two ranks on one GPU: render both shards, add, resolve (all on the renderer's explicit stream)
    with gs.use_stream():
        a = gs.render_shard(0, 3)
        b = gs.render_shard(3, 3)
        total = a + b
    assert np.allclose(total.cpu().numpy().reshape(96, 96, 3), s, rtol=1e-5, atol=1e-6)
    img = gs.resolve(total)
    assert np.abs(img.astype(int) - rgb.astype(int)).max() <= 1
    img1, _ = render_sharded(gs.render_shard, gs.resolve, 6, 0, 1)
    assert np.array_equal(img1, rgb)


@pytest.mark.parametrize("name,w,h", [("cornell_box", 64, 64), ("random_spheres", 96, 54), ("suzanne", 96, 54), ("teapot", 96, 54),
                                      ("hdri_test", 100, 50), ("earth", 64, 64), ("part2_all", 96, 54)])
def test_converged_image_gate(scenes, name, w, h):
    """Gate 3 at reduced resolution (camera framing is resolution independent, camera.rs:89-95): 4096 spp, every
    BASELINE.json config."""
    ns, orc = scenes(name)
    p = params_for(name, w, h, 4096, seed=2)
    grgb, gsum, _ = ns.render(p)
    orgb, osum, _ = orc.render(p)
    assert psnr_u8(grgb, orgb) >= 40.0
    gm, om = gsum / 4096.0, osum / 4096.0
    ok = np.isfinite(om) & np.isfinite(gm)          # part2_all: a NaN-direction path makes a pixel NaN on both sides
    assert np.array_equal(np.isfinite(om), np.isfinite(gm)) and ok.mean() > 0.999
    assert abs(gm[ok].mean() - om[ok].mean()) / abs(om[ok].mean()) < 0.01
    mre = np.mean(np.abs(gm[ok] - om[ok]) / np.maximum(np.abs(om[ok]), 1e-2))
    assert mre < 0.01, mre
    # statistically independent check: a different seed on the GPU must still agree in the mean (Monte-Carlo noise
    # of the no-NEE estimator sets the tolerance: small bright lights converge slowly)
    _, gsum2, _ = ns.render(params_for(name, w, h, 4096, seed=77))
    g2 = gsum2 / 4096.0
    ok2 = ok & np.isfinite(g2)
    tol = 0.02 if name in ("cornell_box", "random_spheres", "earth", "hdri_test") else 0.05
    assert abs(g2[ok2].mean() - om[ok2].mean()) / abs(om[ok2].mean()) < tol


# part2_final.png (examples/part2_all.rs, 600x800): PSNR on 8x8 box means inside the fixed objects (measured: see the print)
#                     brown  earth  noise(turbulence)  metal  glass_small  glass
PART2_PSNR_MIN = {2: 28.0, 7: 28.0, 8: 31.0, 4: 24.0, 3: 20.0, 5: 16.0}
PART2_NAMES = {1: "light", 2: "brown LambertianMat", 3: "small DielectricMat sphere", 4: "MetalMat roughness 10", 5: "DielectricMat + ConstantMedium",
               7: "earth ImageTexture", 8: "TurbulenceTexture(5, 10)"}


def test_part2_final_png_pins_textures_media_and_metal():
    """The reference's committed render part2_final.png is the output of examples/part2_all.rs (camera part2_all.rs:88-91,
    600x800).  Its box heights and small-sphere positions come from an RNG crate that is not vendored, but the light, the
    five large spheres and their materials are constants of the example — and they are exactly the parts of the path no
    other reference artefact exercises: TurbulenceTexture / Perlin noise (the black-patch pattern is position-exact),
    ImageTexture uv on a sphere, DielectricMat, ConstantMedium + IsotropicMat inside glass, MetalMat with roughness.
    The GPU render of scenes/part2_all.yml at the same camera and size is compared with the PNG on 8x8 box means inside
    each fixed object (mask: boxes whose primary rays all hit that object; tests/golden/make_part2_pin.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "part2_final_pin.npz"))
    w, h = (int(v) for v in g["size"])
    ns = native_scene("part2_all")
    rgb, _, _ = ns.render(params_for("part2_all", w, h, 4096, seed=5), want_sum=False)
    ns.close()
    low = rgb.astype(np.float64).reshape(h // 8, 8, w // 8, 8, 3).mean((1, 3))
    d = low - g["low"].astype(np.float64)
    label = g["label"]
    assert (label == 1).sum() > 50 and np.abs(d[label == 1]).max() < 1.0        # the light saturates in both
    for idx, floor in PART2_PSNR_MIN.items():
        m = label == idx
        psnr = 10.0 * np.log10(255.0 ** 2 / np.mean(d[m] ** 2))
        print(f"part2_final.png, {PART2_NAMES[idx]}: {int(m.sum())} boxes, PSNR {psnr:.2f} dB, mean ours {low[m].mean(0).round(1)} "
              f"reference {g['low'][m].mean(0).round(1)}")
        assert psnr >= floor, (PART2_NAMES[idx], psnr)
    # the turbulence pattern is position-exact: shifting our render by one box must be clearly worse than the aligned comparison
    m = label == 8
    aligned = np.mean(d[m] ** 2)
    shifted = np.mean((np.roll(low, 1, axis=1) - g["low"])[m] ** 2)
    assert shifted > 2.0 * aligned, (aligned, shifted)


def _fp32_sums_agree(gsum, osum, min_frac=0.995):
    ok = np.isfinite(osum)
    assert np.array_equal(ok, np.isfinite(gsum))
    close = np.abs(gsum - osum) <= 1e-4 * np.maximum(np.abs(osum), 1e-3)
    close = np.all(close | ~ok, axis=2)
    return close.mean(), close


@pytest.mark.parametrize("name,w,h,spp,batch", [
    ("suzanne", 1920, 1080, 2, 0),        # C3 at BASELINE.json's resolution (1080p)
    ("teapot", 1920, 1080, 2, 0),
    ("part2_all", 3840, 2160, 1, 0),      # C5 at 4K
    ("part2_all", 3840, 2160, 2, 1 << 21),  # ... with batches smaller than the image: pixel tiling (api.cu render_into)
    ("random_spheres", 960, 540, 4, 100_000),
])
def test_full_size_render_matches_oracle(name, w, h, spp, batch):
    """BASELINE.json's own resolutions against the oracle, pixel by pixel (seconds of CPU at 1-2 spp): the 32-bit
    path -> (pixel, sample) mapping (npix_magic), the segment geometry of full-size batches and the pixel-tiled batch
    loop are only exercised at these sizes."""
    ns, orc = native_scene(name), oracle_scene(name, fast=True)
    if batch:
        ns.set_batch_paths(batch)
    p = params_for(name, w, h, spp, seed=9)
    _, gsum, st = ns.render(p, want_rgb=False)
    _, osum, ost = orc.render(p, want_rgb=False)
    ns.close()
    assert st["samples"] == w * h * spp
    frac, close = _fp32_sums_agree(gsum, osum)
    print(f"{name} {w}x{h}x{spp} batch {batch or 'default'}: {100 * frac:.3f} % of pixels within 1e-4 of the oracle, rays {st['rays']} vs {ost['rays']}")
    assert frac >= 0.995, frac
    assert abs(st["rays"] - ost["rays"]) <= 1e-3 * ost["rays"]
    # the few differing pixels are isolated discrete flips (checker sign, Schlick coin, texel, free path), not a region
    bad = ~close[:h // 8 * 8, :w // 8 * 8]
    assert bad.reshape(h // 8, 8, w // 8, 8).sum((1, 3)).max() <= 24


@pytest.mark.parametrize("name,spp", [("random_spheres", 32), ("cornell_box", 1024), ("part2_all", 2)])
def test_full_size_properties(scenes, name, spp):
    """BASELINE.json's own sizes (C1 960x540x32, C2 300x300x1024, C5 at 3840x2160): properties that do not need
    the oracle at that size — determinism, shard additivity, finite non-negative-ish sums, ray accounting."""
    ns, _ = scenes(name)
    cfg = CONFIGS[name]
    p = params_for(name, cfg.width, cfg.height, spp, seed=1)
    rgb, s, st = ns.render(p)
    assert st["samples"] == cfg.width * cfg.height * spp
    assert st["samples"] <= st["rays"] <= 11 * st["samples"]          # depth cap 10 (render.rs:21)
    assert np.isfinite(s).all() or name == "part2_all"                 # turbulence albedo may be negative, not NaN
    half = spp // 2
    _, a, _ = ns.render(params_for(name, cfg.width, cfg.height, spp, seed=1, sample_begin=0, sample_count=half))
    _, b, _ = ns.render(params_for(name, cfg.width, cfg.height, spp, seed=1, sample_begin=half, sample_count=spp - half))
    if spp == 2:
        assert np.array_equal(a + b, s)
    else:
        assert np.allclose(a + b, s, rtol=2e-5, atol=1e-5)
    rgb2, s2, _ = ns.render(p)
    assert np.array_equal(s, s2) and np.array_equal(rgb, rgb2)
    assert rgb.shape == (cfg.height, cfg.width, 3) and rgb.std() > 5.0


def test_cli_checkpoint_resume(tmp_path, scenes):
    """src/main.rs equivalent + sample-range checkpointing: an interrupted render resumed from its checkpoint
    equals the uninterrupted one (same samples; the per-pixel fp32 add of two partial sums is the only difference)."""
    from firework_b200.__main__ import main
    from firework_b200.progressive import render_progressive
    from firework_b200.api import Scene
    from PIL import Image
    scene_file = CONFIGS["conics_cli"].path()
    out = tmp_path / "a.png"
    ck = tmp_path / "ck.npz"
    # 6 of 10 samples, then "crash"; resume to 10
    assert main(["--scene-file", scene_file, "-s", "6", "-o", str(out), "--checkpoint", str(ck), "--chunk", "3",
                 "--width", "160", "--height", "90"]) == 0
    assert int(np.load(ck)["done"]) == 6
    assert main(["--scene-file", scene_file, "-s", "10", "-o", str(out), "--checkpoint", str(ck), "--chunk", "4",
                 "--width", "160", "--height", "90"]) == 0
    assert int(np.load(ck)["done"]) == 10
    resumed = np.asarray(Image.open(out))
    ns, _ = scenes("conics_cli")
    whole, s, _ = ns.render(params_for("conics_cli", 160, 90, 10, seed=0))
    assert resumed.shape == whole.shape
    assert np.abs(resumed.astype(int) - whole.astype(int)).max() <= 1
    assert np.allclose(np.load(ck)["sums"], s, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,spp,gamma", [("suzanne", 512, 2.2), ("teapot", 256, 2.2), ("cornell_box", 1000, 2.0),
                                            ("conics", 256, 2.2), ("earth", 256, 2.2), ("heightmap", 256, 2.2)])
def test_render_matches_the_references_committed_png(scenes, name, spp, gamma):
    """The only end-to-end artefacts the reference ships: renders of scenes that can be reproduced exactly —
    suzanne.png (examples/suzanne.rs:83-96, 960x540), teapot.png (examples/teapot.rs:96-109, 1920x1080) and conics.png
    (examples/conics.rs:85-93) from the committed scenes/*.yml, heightmap.png (examples/heightmap.rs, a procedural
    height-field mesh), cornell_box.png (examples/cornell_box.rs, 300x300;
    rendered by an older revision whose output gamma was 2.0) and Earth.png (examples/earth.rs, 800x800).  The GPU render
    of the same scene at the same resolution must agree with them after an 8x8 box filter (their sample count and RNG
    differ, so the comparison is noise-limited): geometry, orientation, every shape's hit routine, flat vs interpolated
    normals, image-texture uv, the sky, the lights and the output transform all show up here.
    Fixture: tests/golden/reference_png_lowres.npz (generator beside it)."""
    import os
    ns, _ = scenes(name)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_png_lowres.npz"))
    w, h = (int(v) for v in g[name + "_size"])
    p = params_for(name, w, h, spp, seed=5)
    p.gamma = gamma
    rgb, _, _ = ns.render(p, want_sum=False)
    low = rgb[:h // 8 * 8, :w // 8 * 8].astype(np.float64).reshape(h // 8, 8, w // 8, 8, 3).mean((1, 3))
    d = low - g[name].astype(np.float64)
    psnr = 10.0 * np.log10(255.0 ** 2 / np.mean(d ** 2))
    print(f"{name}: PSNR vs the reference's PNG (8x8 box means) {psnr:.2f} dB, mean abs diff {np.abs(d).mean():.2f} / 255")
    assert psnr >= PNG_PSNR_MIN[name], psnr
    assert np.abs(d).mean() <= PNG_MEAN_ABS_MAX[name]


def test_volume_png_background_pins_camera_sky_floor_and_output_transform(scenes):
    """volume.png was rendered by an earlier revision of examples/volume_test.rs (an r = 1.5 medium sphere inside a concentric
    glass shell — silhouette, light reflection and horizon fit that; its colour / density do not, profiles/r02_variants.md §5),
    so the sphere is not comparable.  Everything the sphere does not influence is the current example's: the sky gradient
    (environment.rs:62-66), the horizon row the camera puts it on (camera.rs:74-116, volume_test.rs:62-64), the far floor lit
    by the sky (LambertianMat), gamma and quantisation (render.rs:184-189).  Those 8x8 boxes must agree to a fraction of a
    level: 74 dB with the oracle at 64 spp."""
    import os
    ns, _ = scenes("volume")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_png_lowres.npz"))
    w, h = (int(v) for v in g["volume_size"])
    assert (w, h) == (960, 540)
    rgb, _, _ = ns.render(params_for("volume", w, h, 128, seed=9), want_sum=False)
    low = rgb[:h // 8 * 8, :w // 8 * 8].astype(np.float64).reshape(h // 8, 8, w // 8, 8, 3).mean((1, 3))
    yy, xx = np.mgrid[0:h // 8, 0:w // 8]
    px, py = xx * 8 + 4, yy * 8 + 4
    background = ((px < 250) | (px > 710)) & (py < 100)          # left and right of the sphere, from the top down past the horizon
    d = (low - g["volume"].astype(np.float64))[background]
    psnr = 10.0 * np.log10(255.0 ** 2 / np.mean(d ** 2))
    print(f"volume.png background ({int(background.sum())} boxes): {psnr:.2f} dB, max |diff| {np.abs(d).max():.2f} / 255")
    assert background.sum() > 700
    assert psnr >= 60.0 and np.abs(d).max() <= 1.5
