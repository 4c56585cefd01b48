"""Small fixed workload for ncu captures: python tools/prof_run.py <scene> [w h spp]"""
import os, sys, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
name = sys.argv[1]
w, h, spp = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (960, 540, 8)
cfg = CONFIGS[name]; p = cfg.path()
text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
ns = NativeScene(text, asset_dir=os.path.join(SCENE_DIR, "assets"))
prm = cfg.renderer(width=w, height=h, samples=spp, seed=2).params()
_, _, st = ns.render(prm, want_sum=False)
print(name, st)
ns.close()
