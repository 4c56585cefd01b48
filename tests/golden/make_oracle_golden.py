"""Writes tests/golden/oracle_golden.npz: the oracle's outputs on small fixed inputs for every scene."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import ALL_SCENES, oracle_scene, params_for
out = {}
for name in ALL_SCENES:
    sc = oracle_scene(name)
    p = params_for(name, 24, 16, 2, seed=7)
    o, d = sc.primary_rays(p, 1)
    h = sc.first_hit(o, d, seed=7)
    _, s, _ = sc.render(p, want_rgb=False)
    out.update({f"{name}/o": o, f"{name}/d": d, f"{name}/obj": h["obj"], f"{name}/prim": h["prim"], f"{name}/t": h["t"],
                f"{name}/sum": s})
np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
print("wrote", len(out), "arrays")
