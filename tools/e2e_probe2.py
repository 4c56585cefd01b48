"""e2e step of bench.py with timings (developer tool): python tools/e2e_probe2.py part2_all 128"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS
name, spp = sys.argv[1], int(sys.argv[2])
cfg = CONFIGS[name]
text = bench.read_scene_text(cfg)
doc, assets = bench.predecode_assets(text)
keep = NativeScene(text, assets=assets)
keep.render(cfg.renderer(samples=spp, seed=1).params(), want_sum=False)
for i in range(4):
    t0 = time.perf_counter()
    s = NativeScene(text, assets=assets)
    t1 = time.perf_counter()
    rgb, _, st = s.render(cfg.renderer(samples=spp, seed=1).params(), want_sum=False)
    t2 = time.perf_counter()
    s.close()
    t3 = time.perf_counter()
    print(f"iter {i}: scene {1e3*(t1-t0):.1f} render {1e3*(t2-t1):.1f} (device {st['ms_device']:.1f}) close {1e3*(t3-t2):.1f}", flush=True)
