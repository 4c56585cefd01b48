import os, sys, gzip, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
import torch
ASSETS = os.path.join(SCENE_DIR, "assets")
name = sys.argv[1] if len(sys.argv) > 1 else "cornell_box"
cfg = CONFIGS[name]; p = cfg.path()
text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
keep = NativeScene(text, asset_dir=ASSETS)           # like bench: one scene stays alive
prm = cfg.renderer(samples=int(sys.argv[2]) if len(sys.argv) > 2 else cfg.samples, seed=1).params()
keep.render(prm, want_sum=False)
for i in range(5):
    t0 = time.perf_counter()
    s = NativeScene(text, asset_dir=ASSETS, commit=False)
    t1 = time.perf_counter()
    s.commit()
    t2 = time.perf_counter()
    rgb, _, st = s.render(prm, want_sum=False)
    t3 = time.perf_counter()
    s.close()
    t4 = time.perf_counter()
    print(f"{name} iter {i}: parse {1e3*(t1-t0):7.2f} commit {1e3*(t2-t1):7.2f} render {1e3*(t3-t2):7.2f} (device {st['ms_device']:.2f}) close {1e3*(t4-t3):7.2f} ms", flush=True)
