"""Developer A/B tool: extend-kernel throughput of workloads under env-var knobs, one process."""
import os, sys, json, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
ASSETS = os.path.join(SCENE_DIR, "assets")
SIZES = {"random_spheres": (960, 540, 16), "teapot": (1920, 1080, 4), "suzanne": (1920, 1080, 4), "part2_all": (1920, 1080, 4), "conics_cli": (960, 540, 16)}
def run(name, env):
    for k in ("FW_EXTEND_MODE", "FW_PERSISTENT_FROM", "FW_REFILL_LANES"):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    w, h, spp = SIZES[name]
    cfg = CONFIGS[name]; p = cfg.path()
    text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
    ns = NativeScene(text, asset_dir=ASSETS); ns.set_profiling(True)
    prm = cfg.renderer(width=w, height=h, samples=spp, seed=2).params()
    ns.render(prm, want_sum=False)
    best = None
    for _ in range(3):
        _, _, st = ns.render(prm, want_sum=False)
        if best is None or st["ms_device"] < best["ms_device"]: best = st
    ns.close()
    print(f"{name:15s} {str(env):75s} device {best['ms_device']:8.2f} ms extend {best['ms_extend']:8.2f} ms  "
          f"{best['samples']/best['ms_device']/1e3:8.1f} Msamples/s extend-only {best['rays']/best['ms_extend']/1e3:8.1f} Mrays/s", flush=True)
if __name__ == "__main__":
    for name in sys.argv[1].split(","):
        run(name, {"FW_EXTEND_MODE": 0})
        for frm in (0, 1):
            for r in (8, 16, 24, 32):
                run(name, {"FW_EXTEND_MODE": 1, "FW_PERSISTENT_FROM": frm, "FW_REFILL_LANES": r})
