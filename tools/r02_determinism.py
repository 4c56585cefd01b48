"""Render a workload twice with the walk kernels and once with the lock-step kernels; report differing pixels."""
import os, sys, gzip, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = CONFIGS[name]; p = cfg.path()
text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
def render():
    ns = NativeScene(text, asset_dir=os.path.join(SCENE_DIR, "assets"))
    _, s, st = ns.render(cfg.renderer(width=w, height=h, samples=spp, seed=2).params(), want_rgb=False)
    ns.close()
    return s, st
if len(sys.argv) > 5:      # child: dump
    s, st = render(); np.save(sys.argv[5], s); print(st["rays"]); sys.exit(0)
a, sa = render(); b, sb = render()
print("walk rays", sa["rays"], sb["rays"])
da = ~((a == b) | (np.isnan(a) & np.isnan(b)))
print("walk vs walk: differing values", int(da.sum()), "pixels", np.argwhere(da.any(2))[:10].tolist())
for yx in np.argwhere(da.any(2))[:5]:
    print("   ", yx, a[tuple(yx)], b[tuple(yx)])
out = "/tmp/lock.npy"
r = subprocess.run([sys.executable, __file__, name, str(w), str(h), str(spp), out], env=dict(os.environ, FW_WALK="0"), capture_output=True, text=True)
print("lock rays", r.stdout.strip(), r.stderr[-300:])
c = np.load(out)
dc = ~((a == c) | (np.isnan(a) & np.isnan(c)))
print("walk vs lock: differing values", int(dc.sum()), "pixels", np.argwhere(dc.any(2))[:10].tolist())
for yx in np.argwhere(dc.any(2))[:5]:
    print("   ", yx, a[tuple(yx)], c[tuple(yx)])
