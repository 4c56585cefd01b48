"""fw_render_multi (one process, N GPUs through the C ABI) on the headline workload: Msamples/s and the time of the
combine step for both reduce modes — one ncclReduce of the fp32 sum buffers, or the fused peer-memory reduce + resolve
kernel.  Usage: python tools/multi_bench.py [workload] [spp]"""
import gzip, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from firework_b200 import _native as N
from firework_b200.engine import NativeScene
from firework_b200.scenes import CONFIGS, SCENE_DIR
name = sys.argv[1] if len(sys.argv) > 1 else "part2_all"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cfg = CONFIGS[name]; p = cfg.path()
text = (gzip.open(p, "rt") if p.endswith(".gz") else open(p)).read()
ndev = N.lib().fw_device_count()
ns = NativeScene(text, asset_dir=os.path.join(SCENE_DIR, "assets"))
prm = cfg.renderer(samples=spp, seed=1).params()
ref = None
out = {"workload": name, "width": cfg.width, "height": cfg.height, "spp_total": spp, "devices": ndev, "runs": []}
for n in [g for g in (1, 2, 4, 8) if g <= ndev]:
    for mode in (("nccl", "peer") if n > 1 else ("nccl",)):
        ns.render_multi(prm, n, reduce=mode, want_sum=False)            # warm-up: replicas, communicators, path state
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            rgb, _, st = ns.render_multi(prm, n, reduce=mode, want_sum=False)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]: best = (dt, st)
        if ref is None: ref = rgb
        run = {"n_gpus": n, "reduce": mode, "wall_ms": 1e3 * best[0], "msamples_per_s": cfg.width * cfg.height * spp / best[0] / 1e6,
               "ms_reduce_resolve": best[1]["ms_reduce"], "ms_device_slowest": best[1]["ms_device"],
               "max_abs_diff_vs_1gpu_u8": int(np.abs(rgb.astype(int) - ref.astype(int)).max())}
        out["runs"].append(run)
        print(json.dumps(run), flush=True)
ns.close()
json.dump(out, open(os.path.join("gpurun_out", f"multi_bench_{name}.json"), "w"), indent=1)
