"""Developer A/B tool: extend-kernel throughput of one workload under env-var knobs, one process."""
import os, sys, json, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
ASSETS = os.path.join(SCENE_DIR, "assets")
def run(name, w, h, spp, env):
    for k, v in env.items():
        os.environ[k] = str(v)
    cfg = CONFIGS[name]; p = cfg.path()
    text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
    ns = NativeScene(text, asset_dir=ASSETS); ns.set_profiling(True)
    prm = cfg.renderer(width=w, height=h, samples=spp, seed=1).params()
    ns.render(prm, want_sum=False)
    best = None
    for _ in range(3):
        _, _, st = ns.render(prm, want_sum=False)
        if best is None or st["ms_device"] < best["ms_device"]: best = st
    ns.close()
    print(f"{name:15s} {env} device {best['ms_device']:8.2f} ms extend {best['ms_extend']:8.2f} ms  "
          f"{best['samples']/best['ms_device']/1e3:8.1f} Msamples/s {best['rays']/best['ms_device']/1e3:8.1f} Mrays/s "
          f"extend-only {best['rays']/best['ms_extend']/1e3:8.1f} Mrays/s", flush=True)
if __name__ == "__main__":
    scenes = sys.argv[1].split(",")
    for name in scenes:
        w, h, spp = (960, 540, 16)
        run(name, w, h, spp, {"FW_EXTEND_MODE": 0})
        for r in (1, 8, 16, 22, 28, 32):
            run(name, w, h, spp, {"FW_EXTEND_MODE": 1, "FW_REFILL_LANES": r})
