import os, sys, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
ASSETS = os.path.join(SCENE_DIR, "assets")
def run(name, w, h, spp, env):
    for k, v in env.items(): os.environ[k] = str(v)
    cfg = CONFIGS[name]; p = cfg.path()
    text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
    ns = NativeScene(text, asset_dir=ASSETS); ns.set_profiling(True)
    prm = cfg.renderer(width=w, height=h, samples=spp, seed=1).params()
    ns.render(prm, want_sum=False)
    best = None
    for _ in range(2):
        _, _, st = ns.render(prm, want_sum=False)
        if best is None or st["ms_device"] < best["ms_device"]: best = st
    ns.close()
    print(f"{name:10s} {w}x{h}x{spp} {env}: device {best['ms_device']:8.2f} ms extend {best['ms_extend']:8.2f}  {best['samples']/best['ms_device']/1e3:8.1f} Msamples/s  extend-only {best['rays']/best['ms_extend']/1e3:8.1f} Mrays/s", flush=True)
for mode in (0, 2):
    for (w, h, spp) in ((320, 180, 64), (640, 360, 16), (960, 540, 8), (1920, 1080, 2), (1920, 1080, 8)):
        run("teapot", w, h, spp, {"FW_EXTEND_MODE": mode})
    run("suzanne", 1920, 1080, 8, {"FW_EXTEND_MODE": mode})
    run("random_spheres", 960, 540, 32, {"FW_EXTEND_MODE": mode})
    run("part2_all", 1920, 1080, 8, {"FW_EXTEND_MODE": mode})
run("teapot", 1920, 1080, 8, {"FW_EXTEND_MODE": 0, "FW_BATCH_PATHS": 1000000})
