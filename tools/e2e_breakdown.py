"""Developer tool: where one e2e step goes (host parse / asset hand-over / commit / render + D2H / destroy), per config.
    python tools/e2e_breakdown.py [config ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from firework_b200 import _native as N
from firework_b200.assets import load_asset
from firework_b200.scenes import CONFIGS, SCENE_DIR
import gzip
ASSETS = os.path.join(SCENE_DIR, "assets")
L = N.lib()
def run(name, reps=8):
    cfg = CONFIGS[name]; p = cfg.path()
    text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read().encode()
    prm = cfg.renderer(width=cfg.width, height=cfg.height, samples=cfg.samples, seed=1).params()
    rgb = np.zeros((cfg.height, cfg.width, 3), np.uint8)
    cache = {}
    acc = np.zeros(5)
    for rep in range(reps + 2):
        t = [time.perf_counter()]
        h = C.c_void_p()
        N.check(L.fw_scene_from_yaml(text, len(text), C.byref(h))); t.append(time.perf_counter())
        for i in range(L.fw_scene_num_assets(h)):
            path = L.fw_scene_asset_path(h, i).decode(); kind = L.fw_scene_asset_kind(h, i)
            if path not in cache:
                a = load_asset(path, "hdr" if kind == 1 else "image", ASSETS, None)
                cache[path] = np.ascontiguousarray(a, np.float32 if kind == 1 else np.uint8)
            a = cache[path]
            if kind == 1: N.check(L.fw_scene_set_hdr(h, i, a.shape[1], a.shape[0], N.ptr(a)))
            else: N.check(L.fw_scene_set_image(h, i, a.shape[1], a.shape[0], N.ptr(a)))
        t.append(time.perf_counter())
        N.check(L.fw_scene_commit(h, 0)); t.append(time.perf_counter())
        st = N.FwStats()
        N.check(L.fw_render(h, C.byref(prm), N.ptr(rgb), None, C.byref(st))); t.append(time.perf_counter())
        L.fw_scene_destroy(h); t.append(time.perf_counter())
        if rep >= 2: acc += np.diff(t) * 1e3
    acc /= reps
    print(f"{name:15s} parse {acc[0]:7.2f}  assets {acc[1]:7.2f}  commit {acc[2]:7.2f}  render+d2h {acc[3]:7.2f} (device {st.ms_device:7.2f})  destroy {acc[4]:7.2f}  total {acc.sum():7.2f} ms"
          f"  asset bytes {sum(a.nbytes for a in cache.values())}", flush=True)
for n in (sys.argv[1:] or ["hdri_test", "random_spheres", "earth", "cornell_box"]):
    run(n)
