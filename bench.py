#!/usr/bin/env python
"""bench.py — Msamples/s (and Mrays/s) of the rendering hot path on N B200s, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--spp S] [--impl ours|reference]

A "step" is one full render of the workload (BASELINE.json configs; default = configs[1], cornell_box
300x300 x 1024 spp, the configuration quoted for 1xB200).  Multi-GPU: every rank renders ALL pixels for its own
slice of the sample range (weak scaling: `spp` samples per GPU, spp*N in total, RNG keyed by the global sample
index), one NCCL reduce(SUM) of the fp32 sum buffers to rank 0, resolve on rank 0.

  value    : whole-job Msamples/s, scene already resident in HBM, device-timed (CUDA events, max over ranks)
  e2e      : same metric through the public call with HOST buffers each step: YAML text -> fw_scene_from_yaml ->
             fw_scene_commit (H2D) -> render -> u8 image in host memory (D2H)
  roofline : extend kernel (closest-hit traversal), algorithmic FP32 flops per ray (SURVEY.md §8d formula with
             per-ray test counts measured by the oracle's counters) / CUDA-event time of the extend launches,
             against the FP32 FMA peak measured on this GPU in the same run; L2 and HBM views beside it
  cpu_baseline : the C++ oracle ("port" — the Rust reference cannot be built here) on all host cores, bounded sample
  --impl reference : the same CPU port as the reference arm (rank 0 only).
"""
import argparse
import gzip
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

from firework_b200.assets import load_asset  # noqa: E402
from firework_b200.scenes import CONFIGS, SCENE_DIR  # noqa: E402

ASSETS = os.path.join(SCENE_DIR, "assets")
DEFAULT_WORKLOAD = "cornell_box"   # BASELINE.json configs[1]

# static per-test operation / byte counts of the reference's routines
F_NODE, F_SPHERE, F_RECT, F_TRI, F_CONIC, F_XROT, F_XTRANS, F_SHADE = 27, 40, 12, 60, 50, 33, 3, 45
B_NODE, B_SPHERE, B_RECT, B_TRI, B_CONIC, B_XROT, B_XTRANS, B_RAY = 32, 16, 32, 48, 16, 112, 16, 64
HBM_BYTES_PER_RAY = 176  # wavefront streams per extend + shade round (queues carry the records, all coalesced):
                         # extend: ray 32 r + hit record 48 w;  shade: hit record 48 r + next ray 32 w + attenuation 16 w


def read_scene_text(cfg):
    p = cfg.path()
    with (gzip.open(p, "rt") if p.endswith(".gz") else open(p)) as f:
        return f.read()


def predecode_assets(text):
    """Decode every asset the document names once, into host memory (the e2e loop then only copies)."""
    from firework_b200.serde_yaml import loads
    doc = loads(text)
    out = {}

    def walk(n):
        if isinstance(n, dict):
            if n.get("texture") == "ImageTexture":
                out[n["value"]] = load_asset(n["value"], "image", ASSETS)
            if n.get("environment") == "HdrEnvironment":
                out[n["value"]] = load_asset(n["value"], "hdr", ASSETS)
            for v in n.values():
                walk(v)
        elif isinstance(n, list):
            for v in n:
                walk(v)

    walk({"materials": doc["materials"], "environment": doc["environment"]})
    return doc, out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_scene_for(doc, cfg, assets):
    from oracle.oracle import OracleScene
    return OracleScene(doc, cfg.use_bvh, asset_loader=lambda p, k: assets[p], fast=True)


def per_ray_work(st):
    """Algorithmic flops / bytes per ray from the oracle's counters (reference traversal semantics)."""
    r = max(st["rays"], 1)
    flops = (st["aabb_tests"] * F_NODE + st["sphere"] * F_SPHERE + st["rect"] * F_RECT + st["tri"] * F_TRI +
             st["conic"] * F_CONIC + st["xform_rot"] * F_XROT + st["xform_trans"] * F_XTRANS) / r + F_SHADE
    byts = (st["aabb_tests"] * B_NODE + st["sphere"] * B_SPHERE + st["rect"] * B_RECT + st["tri"] * B_TRI +
            st["conic"] * B_CONIC + st["xform_rot"] * B_XROT + st["xform_trans"] * B_XTRANS) / r + B_RAY
    return flops, byts, {"node_tests_per_ray": st["aabb_tests"] / r, "prim_tests_per_ray": st["prim_tests"] / r}


def run_cpu(cfg, doc, assets, width, height, spp, target_seconds, seed):
    """The CPU port on all host cores over a bounded sample of the workload (same scene, same resolution,
    reduced spp where needed). Returns (Msamples/s, Mrays/s, stats, sample description)."""
    orc = oracle_scene_for(doc, cfg, assets)
    r = cfg.renderer(width=width, height=height, samples=spp, seed=seed)
    threads = len(os.sched_getaffinity(0))   # all host cores (torchrun exports OMP_NUM_THREADS=1: override it)
    probe_spp = 1
    _, _, st = orc.render(r.params(sample_begin=0, sample_count=probe_spp), want_rgb=False, threads=threads)
    rate = st["samples"] / max(st["seconds"], 1e-6)
    n = int(max(1, min(spp, target_seconds * rate / (width * height))))
    _, _, st = orc.render(r.params(sample_begin=0, sample_count=n), want_rgb=False, threads=threads)
    return st["samples"] / st["seconds"] / 1e6, st["rays"] / st["seconds"] / 1e6, st, f"{width}x{height} x {n} of {spp} spp"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(CONFIGS))
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel PER GPU (default: the config's)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    cfg = CONFIGS[args.workload]
    width, height = args.width or cfg.width, args.height or cfg.height
    spp = args.spp or cfg.samples
    spp_total = spp * world
    text = read_scene_text(cfg)
    doc, assets = predecode_assets(text)
    config = {"workload": f"{cfg.name} ({cfg.cite})", "width": width, "height": height, "spp_per_gpu": spp,
              "spp_total": spp_total, "use_bvh": cfg.use_bvh, "sharding": f"sample range split over {world} rank(s)",
              "l2": "256 MiB device write between timed steps; per-batch path state (>1 GB) exceeds L2"}
    base = {"metric": "Msamples/s", "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (reference example scene restated; seeded generator where the example is random)",
            "config": config}

    # ------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        # The reference is Rust and cannot be built in this image: the arm times the C++ port of its CPU path
        # (oracle/), all host threads, each step a bounded sample of the same workload.
        vals, rays, st, sample = [], [], None, ""
        for i in range(args.warmup + args.steps):
            ms, mr, st, sample = run_cpu(cfg, doc, assets, width, height, spp, max(2.0, args.cpu_seconds / 2), seed=i)
            if i >= args.warmup:
                vals.append(ms); rays.append(mr)
        v = float(np.mean(vals))
        line = dict(base)
        line.update({"impl": "reference", "value": v, "mrays_per_s": float(np.mean(rays)),
                     "ms_per_step": 1e3 * (width * height * spp_total) / (v * 1e6),
                     "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": st["threads"], "kind": "port", "sample": sample},
                     "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from firework_b200.distributed import GpuShardRenderer, render_sharded, shard_range
    from firework_b200.engine import NativeScene, measure_peaks

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = measure_peaks(local_rank)
    ns = NativeScene(text, device=local_rank, assets=assets)
    ns.set_profiling(True)
    renderer = cfg.renderer(width=width, height=height, samples=spp_total, seed=1)
    shard = GpuShardRenderer(ns, renderer, local_rank)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    npix = width * height

    totals = {"rays": 0, "launches": 0, "ms_extend": 0.0, "extend_launches": 0}

    def render_shard(begin, count):
        t = shard.render_shard(begin, count)
        if shard.last_stats and count > 0:
            st = shard.last_stats
            totals["rays"] += st["rays"]; totals["launches"] += st["launches"] + 1  # + resolve / zero-fill
            totals["ms_extend"] += st["ms_extend"]; totals["extend_launches"] += st["extend_launches"]
        return t

    def step():
        img, _ = render_sharded(render_shard, shard.resolve, spp_total, rank, world)
        return img

    # nvidia-smi needs ~0.1-0.3 s before its first sample: start it before the warm-up so that even a 150 ms timed
    # region is covered; only samples inside the timed window are used
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(warmup):
        step()
    barrier()
    for k in totals:
        totals[k] = 0
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    for e0, e1 in evs:
        flush_buf.fill_(1)          # evict L2 between timed steps (not timed)
        barrier()
        e0.record()
        step()
        e1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_total = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    tt = torch.tensor([ms_total, float(totals["rays"]), float(totals["launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, rays_all, launches_all = float(mx[0]), float(sm[1]), int(sm[2])
    else:
        rays_all, launches_all = float(tt[1]), int(tt[2])
    samples_all = float(npix) * spp_total * args.steps
    value = samples_all / (ms_total * 1e-3) / 1e6
    mrays = rays_all / (ms_total * 1e-3) / 1e6

    # ---- e2e: host YAML text + host texels in, host u8 image out, every step ---------------------------------
    def e2e_step():
        t_a = time.perf_counter()
        s = NativeScene(text, device=local_rank, assets=assets)      # parse, BVH build, flatten, H2D
        t_b = time.perf_counter()
        r2 = cfg.renderer(width=width, height=height, samples=spp_total, seed=1)
        if world == 1:
            rgb, _, st2 = s.render(r2.params(), want_sum=False)      # render + D2H into a host buffer
            d2h = rgb.nbytes
            if os.environ.get("FW_BENCH_DEBUG"):
                print(f"[bench] e2e scene {1e3 * (t_b - t_a):.1f} ms, render call {1e3 * (time.perf_counter() - t_b):.1f} ms, "
                      f"device {st2['ms_device']:.1f} ms", file=sys.stderr)
        else:
            sh = GpuShardRenderer(s, r2, local_rank)
            rgb, _ = render_sharded(sh.render_shard, sh.resolve, spp_total, rank, world)
            d2h = rgb.nbytes if rgb is not None else 0
        h2d = s.device_bytes()
        s.close()
        return h2d, d2h

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        t_step = time.perf_counter()
        a, b = e2e_step()
        h2d, d2h = a, max(d2h, b)
        if os.environ.get("FW_BENCH_DEBUG"):
            print(f"[bench] e2e step {1e3 * (time.perf_counter() - t_step):.1f} ms", file=sys.stderr)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(npix) * spp_total * n_e2e / float(te[0]) / 1e6

    if rank == 0:
        # per-ray algorithmic work from the oracle's counters: one 1-spp pass of the same workload
        orc = oracle_scene_for(doc, cfg, assets)
        _, _, cst = orc.render(cfg.renderer(width=width, height=height, samples=1, seed=1).params(), want_rgb=False,
                               threads=len(os.sched_getaffinity(0)))
        flops_ray, bytes_ray, per_ray = per_ray_work(cst)
        ext_s = max(totals["ms_extend"], 1e-9) * 1e-3
        my_rays = totals["rays"]
        ach_tflops = flops_ray * my_rays / ext_s / 1e12
        ach_l2 = bytes_ray * my_rays / ext_s / 1e9
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of this very
        # command (profiles/r01_extend_traffic_<workload>.json, written by tools/ncu_traffic.py); null if absent
        traffic = None
        try:
            tj = json.load(open(os.path.join(REPO, "profiles", f"r01_extend_traffic_{cfg.name}.json")))
            if spp == cfg.samples and width == cfg.width and height == cfg.height:
                traffic = tj["traffic_bytes_per_launch"]
        except Exception:
            pass
        line = dict(base)
        line.update({
            "value": value, "mrays_per_s": mrays, "ms_per_step": ms_total / args.steps,
            "rays_per_sample": rays_all / samples_all,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": n_e2e, "what": "YAML text + host texels -> fw_scene_from_yaml + commit + render -> host u8 image"},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "kernel": "extend_kernel", "achieved": ach_tflops, "peak": peaks["fp32_tflops"],
                         "unit": "TFLOP/s", "frac": ach_tflops / peaks["fp32_tflops"], "traffic": traffic,
                         "peak_source": "FP32 FMA microbenchmark on this GPU in this run (fw_measure_peaks); MEASURED_PEAKS.json has no FP32 figure",
                         "flops_per_ray": flops_ray, "rays_per_launch": my_rays / max(totals["extend_launches"], 1),
                         "avg_launch_ms": totals["ms_extend"] / max(totals["extend_launches"], 1),
                         "extend_share_of_step": totals["ms_extend"] / (ms_total if world == 1 else max(ms_total, 1e-9)),
                         **per_ray},
            "roofline_l2": {"bound": "l2", "achieved": ach_l2, "peak": peaks["l2_gbs"], "unit": "GB/s",
                            "frac": ach_l2 / peaks["l2_gbs"], "bytes_per_ray": bytes_ray,
                            "peak_source": "L2-resident 48 MiB streaming read on this GPU in this run"},
            "roofline_hbm": {"bound": "hbm", "achieved": HBM_BYTES_PER_RAY * rays_all / (ms_total * 1e-3) / 1e9 / world,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": HBM_BYTES_PER_RAY * rays_all / (ms_total * 1e-3) / 1e9 / world / hbm_peak,
                             "bytes_per_ray": HBM_BYTES_PER_RAY, "note": "wavefront state/queue streams; scene data is L2-resident"},
            "peaks_measured": peaks,
        })
        if world == 1 and not args.no_cpu_baseline:
            ms_cpu, mr_cpu, st, sample = run_cpu(cfg, doc, assets, width, height, spp, args.cpu_seconds, seed=1)
            line["cpu_baseline"] = {"value": ms_cpu, "unit": "Msamples/s", "mrays_per_s": mr_cpu, "cores": st["threads"],
                                    "kind": "port", "sample": sample}
        print(json.dumps(line))
    ns.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
