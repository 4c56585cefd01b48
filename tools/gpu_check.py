"""Developer smoke run on a GPU box: parity of every stage against the oracle + quick timings.
Usage: python tools/gpu_check.py [scene ...]      (writes a summary to stdout)
"""
import gzip
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firework_b200.assets import load_asset  # noqa: E402
from firework_b200.engine import NativeScene, measure_peaks  # noqa: E402
from firework_b200.scenes import CONFIGS, SCENE_DIR  # noqa: E402
from firework_b200.serde_yaml import loads  # noqa: E402
from oracle.oracle import OracleScene  # noqa: E402

ASSETS = os.path.join(SCENE_DIR, "assets")


def read(cfg):
    p = cfg.path()
    return (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-20)


def check(name, w, h, spp, time_spp):
    cfg = CONFIGS[name]
    text = read(cfg)
    t0 = time.time()
    ns = NativeScene(text, asset_dir=ASSETS)
    t_native = time.time() - t0
    orc = OracleScene(loads(text), cfg.use_bvh, asset_loader=lambda p, k: load_asset(p, k, ASSETS), fast=True)
    r = cfg.renderer(width=w, height=h, samples=spp, seed=1234)
    p = r.params()
    # 1. primary rays
    go, gd = ns.primary_rays(p, 0)
    oo, od = orc.primary_rays(p, 0)
    prim_ok = np.array_equal(go, oo) and np.array_equal(gd, od)
    # 2. first hit on primary rays + a second generation of rays leaving the hit points
    res = {}
    for gen in range(3):
        n = len(go)
        pix = np.arange(n, dtype=np.uint32)
        smp = np.zeros(n, np.uint32)
        bnc = np.full(n, gen, np.uint32)
        g = ns.first_hit(go, gd, cfg.use_bvh, seed=1234, pixel=pix, sample=smp, bounce=bnc)
        o = orc.first_hit(go, gd, seed=1234, pixel=pix, sample=smp, bounce=bnc)
        hit = o["obj"] >= 0
        ids_equal = np.array_equal(g["obj"], o["obj"]) and np.array_equal(g["prim"], o["prim"])
        nbad = int(np.sum((g["obj"] != o["obj"]) | (g["prim"] != o["prim"])))
        t_bits = int(np.sum(g["t"][hit] != o["t"][hit]))
        t_rel = float(rel_err(g["t"][hit], o["t"][hit]).max()) if hit.any() else 0.0
        both = hit & (g["obj"] == o["obj"])
        n_rel = float(np.nanmax(np.abs(g["normal"][both] - o["normal"][both]) /
                                np.maximum(np.linalg.norm(o["normal"][both], axis=1, keepdims=True), 1e-20))) if both.any() else 0.0
        uv_abs = float(np.nanmax(np.abs(g["uv"][both] - o["uv"][both]))) if both.any() else 0.0
        res[gen] = dict(ids_equal=ids_equal, nbad=nbad, hits=int(hit.sum()), t_not_biteq=t_bits, t_rel=t_rel,
                        n_rel=n_rel, uv_abs=uv_abs, gpu_nodes_per_ray=g["node_tests"] / n,
                        orc_nodes_per_ray=o["aabb_tests"] / n, gpu_prims_per_ray=g["prim_tests"] / n,
                        orc_prims_per_ray=o["prim_tests"] / n)
        # next generation: leave the oracle's hit point along (normal + fixed pseudo-random offset)
        rng = np.random.default_rng(gen)
        off = rng.uniform(-0.7, 0.7, size=(n, 3)).astype(np.float32)
        nrm = o["normal"] / np.maximum(np.linalg.norm(o["normal"], axis=1, keepdims=True), 1e-20)
        go = np.where(hit[:, None], o["point"], go).astype(np.float32)
        gd = np.where(hit[:, None], (nrm + off), gd).astype(np.float32)
    # 3. low-spp render parity
    grgb, gsum, gst = ns.render(p)
    orgb, osum, ost = orc.render(p)
    diff = np.abs(gsum - osum)
    pix_bad = int(np.sum(np.any(diff > 1e-4 * np.maximum(np.abs(osum), 1e-3), axis=2)))
    rgb_bad = int(np.sum(np.any(grgb != orgb, axis=2)))
    rgb_max = int(np.abs(grgb.astype(int) - orgb.astype(int)).max())
    mean_rel = float(np.abs(gsum.mean() - osum.mean()) / max(abs(osum.mean()), 1e-9))
    # 4. timing
    rt = cfg.renderer(width=w, height=h, samples=time_spp, seed=1)
    pt = rt.params()
    ns.render(pt, want_sum=False)
    ns.set_profiling(True)
    t0 = time.time()
    _, _, st = ns.render(pt, want_sum=False)
    wall = time.time() - t0
    print(f"== {name} {w}x{h}: native load+commit {t_native*1e3:.1f} ms | primary rays bit-equal: {prim_ok}")
    for gen, rr in res.items():
        print(f"   first-hit gen{gen}: {rr}")
    print(f"   render {spp} spp: pixels with sum mismatch {pix_bad}/{w*h}, u8 mismatch {rgb_bad} (max diff {rgb_max}), "
          f"mean rel diff {mean_rel:.3e}, gpu rays {gst['rays']} oracle rays {ost['rays']}")
    print(f"   timing {time_spp} spp: device {st['ms_device']:.2f} ms (wall {wall*1e3:.2f}), extend {st['ms_extend']:.2f} ms, "
          f"{st['samples']/st['ms_device']/1e3:.1f} Msamples/s, {st['rays']/st['ms_device']/1e3:.1f} Mrays/s, "
          f"launches {st['launches']}; oracle {ost['samples']/ost['seconds']/1e6:.2f} Msamples/s on {ost['threads']} threads")
    sys.stdout.flush()
    ns.close()


if __name__ == "__main__":
    print("peaks:", measure_peaks(0))
    names = sys.argv[1:] or ["random_spheres", "cornell_box", "suzanne", "teapot", "earth", "hdri_test", "conics",
                             "conics_cli", "volume", "part2_all"]
    for nme in names:
        try:
            check(nme, 320, 180, 4, 64)
        except Exception as e:  # keep going: this is a survey run
            import traceback
            traceback.print_exc()
            print(f"== {nme}: FAILED {e}")
