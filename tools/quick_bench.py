"""Developer tool: device-timed render of several workloads in one process (best of 3)."""
import os, sys, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
ASSETS = os.path.join(SCENE_DIR, "assets")
DEFAULT = {"random_spheres": (960, 540, 32), "cornell_box": (300, 300, 256), "suzanne": (1920, 1080, 16), "teapot": (1920, 1080, 16),
           "part2_all": (3840, 2160, 8), "earth": (800, 800, 32), "hdri_test": (500, 250, 128), "conics_cli": (960, 540, 32), "volume": (960, 540, 32)}
def run(name):
    cfg = CONFIGS[name]; p = cfg.path()
    w, h, spp = DEFAULT[name]
    text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
    ns = NativeScene(text, asset_dir=ASSETS); ns.set_profiling(True)
    prm = cfg.renderer(width=w, height=h, samples=spp, seed=1).params()
    ns.render(prm, want_sum=False)
    best = None
    for _ in range(3):
        _, _, st = ns.render(prm, want_sum=False)
        if best is None or st["ms_device"] < best["ms_device"]: best = st
    ns.close()
    print(f"{name:15s} {w}x{h}x{spp}: device {best['ms_device']:8.2f} ms extend {best['ms_extend']:8.2f} ms  "
          f"{best['samples']/best['ms_device']/1e3:8.1f} Msamples/s {best['rays']/best['ms_device']/1e3:8.1f} Mrays/s "
          f"extend-only {best['rays']/best['ms_extend']/1e3:8.1f} Mrays/s", flush=True)
if __name__ == "__main__":
    for n in (sys.argv[1:] or list(DEFAULT)):
        run(n)
