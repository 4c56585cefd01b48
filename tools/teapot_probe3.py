import os, sys, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
ASSETS = os.path.join(SCENE_DIR, "assets")
cfg = CONFIGS["teapot"]; p = cfg.path()
text = gzip.open(p, "rt").read()
ns = NativeScene(text, asset_dir=ASSETS); ns.set_profiling(True)
r = cfg.renderer(width=1920, height=1080, samples=8, seed=1)
os.environ["FW_DEBUG_STEPS"] = "1"
for s0 in (5, 6):
    _, _, st = ns.render(r.params(sample_begin=s0, sample_count=1), want_sum=False)
    print("sample", s0, f"device {st['ms_device']:.2f} extend {st['ms_extend']:.2f} rays {st['rays']}", flush=True)
