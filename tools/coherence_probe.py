import os, sys, gzip, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
ASSETS = os.path.join(SCENE_DIR, "assets")
def simt(steps):
    n = len(steps) // 32 * 32
    g = steps[:n].reshape(-1, 32)
    return g.sum() / (32.0 * np.maximum(g.max(1), 1).sum())
for name, (w, h) in {"random_spheres": (960, 540), "teapot": (960, 540), "part2_all": (960, 540)}.items():
    cfg = CONFIGS[name]; p = cfg.path()
    text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
    ns = NativeScene(text, asset_dir=ASSETS)
    os.environ["FW_DEBUG_STEPS"] = "1"; os.environ["FW_DEBUG_DUMP"] = f"/tmp/dump_{name}"
    ns.render(cfg.renderer(width=w, height=h, samples=2, seed=2).params(), want_sum=False)
    del os.environ["FW_DEBUG_STEPS"]; del os.environ["FW_DEBUG_DUMP"]
    ns.close()
    for b in range(0, 5):
        raw = np.fromfile(f"/tmp/dump_{name}_b{b}.bin", dtype=np.float32)
        cnt = raw[:1].view(np.uint32)[0]
        rec = raw[1:].reshape(-1, 8)
        o, d, steps = rec[:, 0:3], rec[:, 3:6], rec[:, 6]
        base = simt(steps)
        octant = (d[:, 0] < 0).astype(int) | ((d[:, 1] < 0).astype(int) << 1) | ((d[:, 2] < 0).astype(int) << 2)
        by_oct = simt(steps[np.argsort(octant, kind="stable")])
        # origin cell (8^3 grid over the bounding box of origins) then octant
        lo, hi = o.min(0), o.max(0)
        cell = np.clip(((o - lo) / np.maximum(hi - lo, 1e-6) * 8).astype(int), 0, 7)
        key = ((cell[:, 0] * 8 + cell[:, 1]) * 8 + cell[:, 2]) * 8 + octant
        by_cell = simt(steps[np.argsort(key, kind="stable")])
        ideal = simt(np.sort(steps))
        print(f"{name:15s} bounce {b}: rays {cnt:8d} mean box tests {steps.mean():6.1f}  SIMT proxy: queue order {base:.3f}  octant {by_oct:.3f}  cell+octant {by_cell:.3f}  sorted-by-cost (upper bound) {ideal:.3f}", flush=True)
