import os, sys, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
ASSETS = os.path.join(SCENE_DIR, "assets")
def run(name, w, h, spp, paths):
    cfg = CONFIGS[name]; p = cfg.path()
    text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
    ns = NativeScene(text, asset_dir=ASSETS); ns.set_batch_paths(paths)
    prm = cfg.renderer(width=w, height=h, samples=spp, seed=1).params()
    ns.render(prm, want_sum=False)
    best = None
    for _ in range(3):
        _, _, st = ns.render(prm, want_sum=False)
        if best is None or st["ms_device"] < best["ms_device"]: best = st
    ns.close()
    print(f"{name:15s} {w}x{h}x{spp} batch {paths:>9d}: device {best['ms_device']:8.2f} ms {best['samples']/best['ms_device']/1e3:8.1f} Msamples/s launches {best['launches']}", flush=True)
for paths in (1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25):
    run("cornell_box", 300, 300, 1024, paths)
for paths in (1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25):
    run("random_spheres", 960, 540, 32, paths)
    run("teapot", 1920, 1080, 16, paths)
