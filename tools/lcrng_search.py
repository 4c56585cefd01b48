"""Which generator is tiny_rng 0.1.0's `Rng` / `LcRng`?  (crate not vendored; Cargo.lock:965-968)

The reference's committed render random_spheres.png was produced by examples/random_spheres.rs from
`Rng::new(12345)`: the positions / materials of its ~480 small spheres are a fingerprint of the generator.  Each candidate
(generator family x constants x output bits x float conversion) regenerates the scene through the mirrored API
(firework_b200/scenes.py random_scene follows examples/random_spheres.rs:14-67 line by line), the CPU oracle renders it at
low spp, and the render is compared with the PNG on box means.  A wrong stream puts the spheres elsewhere (~17-19 dB);
the right one reproduces the image up to Monte-Carlo noise.

    python tools/lcrng_search.py            # run in the build container (reads /root/reference/random_spheres.png)
"""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image
from firework_b200.api import F
from firework_b200.scenes import CONFIGS, final_scene, random_scene
from firework_b200.serde_yaml import loads
from oracle.oracle import OracleScene

M64 = (1 << 64) - 1
REF = np.asarray(Image.open("/root/reference/random_spheres.png").convert("RGB")).astype(np.float64)
REF2 = np.asarray(Image.open("/root/reference/part2_final.png").convert("RGB")).astype(np.float64)   # 600 x 800


REF2_LUM = REF2.reshape(400, 2, 300, 2, 3).mean((1, 3, 4))     # 300 x 400 luminance
_rays = {}


def score_part2(rng, spp=0):
    """examples/part2_all.rs: the generator places the 1000 small white spheres (part2_all.rs:62-66); their cloud's bumpy
    silhouette against the dark background is a sharp, noise-free fingerprint.  Primary rays only: the oracle's first-hit
    object ids give our silhouette, thresholded luminance gives the PNG's; score = intersection over union."""
    scene = final_scene(rng)
    doc = loads(scene.to_yaml())
    from firework_b200.assets import load_asset
    from firework_b200.scenes import SCENE_DIR
    orc = OracleScene(doc, True, fast=True, asset_loader=lambda p_, k_: load_asset(p_, k_, os.path.join(SCENE_DIR, "assets")))
    cfg = CONFIGS["part2_all"]
    w, h = 300, 400
    if "od" not in _rays:
        _rays["od"] = orc.primary_rays(cfg.renderer(width=w, height=h, samples=1, seed=1).params(), 0)
    o, d = _rays["od"]
    hit = orc.first_hit(o, d, seed=1)
    obj = hit["obj"].reshape(h, w)
    n = len(doc["render_objects"])
    cloud = (obj >= n - 1001) & (obj < n - 1)          # the 1000 spheres are the last objects before the global medium
    # the global medium (last object) may win the first hit in front of a sphere only rarely (density 1e-4): ignore
    ys, xs = slice(100, 245), slice(140, 300)
    ref_mask = REF2_LUM[ys, xs] > 38.0
    ours = cloud[ys, xs]
    iou = (ours & ref_mask).sum() / max((ours | ref_mask).sum(), 1)
    return 100 * iou, 100 * iou, n


class Lcg:
    def __init__(self, seed, a, c, bits, out, conv, init):
        self.a, self.c, self.mask, self.out, self.conv = a, c, (1 << bits) - 1, out, conv
        s = seed & self.mask
        if init == "seed+c": s = (s + c) & self.mask
        elif init == "step": s = (s * a + c) & self.mask
        elif init == "pcg": s = ((((0 * a + c) & self.mask) + seed) * a + c) & self.mask
        self.s = s
        self.bits = bits

    def u32(self):
        self.s = (self.s * self.a + self.c) & self.mask
        s = self.s
        if self.bits == 64:
            if self.out == "hi": return s >> 32
            if self.out == "lo": return s & 0xFFFFFFFF
            if self.out == "mid": return (s >> 16) & 0xFFFFFFFF
            if self.out == "xsh": return ((s ^ (s >> 22)) >> (22 + (s >> 61))) & 0xFFFFFFFF       # PCG RXS-M-XS-ish
            if self.out == "xshrr":
                x = (((s >> 18) ^ s) >> 27) & 0xFFFFFFFF; r = s >> 59
                return ((x >> r) | (x << ((-r) & 31))) & 0xFFFFFFFF
        return s & 0xFFFFFFFF

    def rand_f32(self):
        return to_f32(self.u32(), self.conv)


class XorShift:
    def __init__(self, seed, kind, conv):
        self.kind, self.conv = kind, conv
        if kind == "xs128":
            self.x, self.y, self.z, self.w = (123456789 ^ seed) & 0xFFFFFFFF, (362436069 ^ seed) & 0xFFFFFFFF, 521288629, 88675123
        elif kind == "xs64":
            self.s = seed & M64
        elif kind == "xs64star":
            self.s = seed & M64
        elif kind == "splitmix":
            self.s = seed & M64

    def u32(self):
        k = self.kind
        if k == "xs128":
            t = (self.x ^ (self.x << 11)) & 0xFFFFFFFF
            self.x, self.y, self.z = self.y, self.z, self.w
            self.w = (self.w ^ (self.w >> 19) ^ t ^ (t >> 8)) & 0xFFFFFFFF
            return self.w
        if k == "xs64":
            s = self.s; s ^= (s << 13) & M64; s ^= s >> 7; s ^= (s << 17) & M64; self.s = s
            return s >> 32
        if k == "xs64star":
            s = self.s; s ^= s >> 12; s ^= (s << 25) & M64; s ^= s >> 27; self.s = s
            return ((s * 0x2545F4914F6CDD1D) & M64) >> 32
        if k == "splitmix":
            self.s = (self.s + 0x9E3779B97F4A7C15) & M64
            z = self.s; z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64; z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
            return (z ^ (z >> 31)) >> 32

    def rand_f32(self):
        return to_f32(self.u32(), self.conv)


def to_f32(u, conv):
    if conv == "2^-32": return F(np.float32(u) * np.float32(2.3283064365386963e-10))
    if conv == "/max": return F(np.float32(u) / np.float32(4294967295.0))
    if conv == ">>8": return F((u >> 8) * (1.0 / 16777216.0))
    if conv == "mant": return F(np.frombuffer(np.uint32(0x3F800000 | (u >> 9)).tobytes(), np.float32)[0] - np.float32(1.0))


def score(rng, w=480, h=270, spp=3):
    scene = random_scene(rng)
    doc = loads(scene.to_yaml())
    orc = OracleScene(doc, True, fast=True)
    cfg = CONFIGS["random_spheres"]
    rgb, _, _ = orc.render(cfg.renderer(width=w, height=h, samples=spp, seed=1).params(), threads=8)
    k = 8
    a = rgb[:h // k * k, :w // k * k].astype(np.float64).reshape(h // k, k, w // k, k, 3).mean((1, 3))
    f = 960 // w * k
    b = REF[:540 // f * f, :960 // f * f].reshape(540 // f, f, 960 // f, f, 3).mean((1, 3))
    d = a - b
    lower = d[d.shape[0] // 2:]            # the ground / small-sphere half of the image
    return 10 * np.log10(255 ** 2 / np.mean(d ** 2)), 10 * np.log10(255 ** 2 / np.mean(lower ** 2)), len(doc["render_objects"])


def candidates():
    mults64 = {"mmix": 6364136223846793005, "lecuyer1": 2862933555777941757, "lecuyer2": 3202034522624059733, "lecuyer3": 3935559000370003845}
    incs = {"mmix": 1442695040888963407, "1": 1, "pcgdef": 1442695040888963407 | 1, "0": 0, "11": 11, "12345": 12345}
    for (an, a), (cn, c), out, conv, init in itertools.product(mults64.items(), incs.items(), ["hi", "lo", "mid", "xsh", "xshrr"],
                                                               ["2^-32", "/max", ">>8", "mant"], ["seed", "seed+c", "step", "pcg"]):
        if cn == "pcgdef" and an != "mmix": continue
        yield f"lcg64 a={an} c={cn} out={out} conv={conv} init={init}", (lambda a=a, c=c, out=out, conv=conv, init=init: Lcg(12345, a, c, 64, out, conv, init))
    for (an, a, c), conv, init in itertools.product([("ansi", 1103515245, 12345), ("nr", 1664525, 1013904223), ("msvc", 214013, 2531011)],
                                                    ["2^-32", "/max", ">>8", "mant"], ["seed", "step"]):
        yield f"lcg32 {an} conv={conv} init={init}", (lambda a=a, c=c, conv=conv, init=init: Lcg(12345, a, c, 32, "lo", conv, init))
    for kind, conv in itertools.product(["xs128", "xs64", "xs64star", "splitmix"], ["2^-32", "/max", ">>8", "mant"]):
        yield f"{kind} conv={conv}", (lambda kind=kind, conv=conv: XorShift(12345, kind, conv))


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    best = []
    t0 = time.time()
    for i, (name, make) in enumerate(candidates()):
        if only and only not in name: continue
        try:
            s_all, s_low, nobj = (score_part2 if os.environ.get("LCRNG_SCENE", "part2") == "part2" else score)(make())
        except Exception as e:   # a stream that hits the example's unreachable!() arm etc.
            print(f"{name}: {e}", flush=True); continue
        best.append((s_low, s_all, name, nobj))
        if s_low > float(os.environ.get('LCRNG_PRINT_ABOVE', '75.0')) or i % 50 == 0:
            print(f"[{i} {time.time() - t0:.0f}s] {name}: all {s_all:.2f} dB, lower half {s_low:.2f} dB, objects {nobj}", flush=True)
    best.sort(reverse=True)
    print("TOP:")
    for b in best[:15]:
        print(f"  lower {b[0]:.2f} dB  all {b[1]:.2f} dB  {b[2]} ({b[3]} objects)")
