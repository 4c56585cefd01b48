"""Reads an .ncu-rep holding `--set full` captures of the extend kernel and writes the per-launch DRAM traffic /
duration summary that bench.py reports as roofline.traffic.  Usage: python tools/ncu_traffic.py in.ncu-rep out.json"""
import csv, io, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
def val(r, name):
    i = col(name); v = float(r[i].replace(",", "")); u = units[i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1)
    return v * scale
launches = []
for r in rows[2:]:
    if not any(k in r[col("Kernel Name")] for k in ("extend", "walk_mesh", "classify_mesh")): continue
    launches.append({"kernel": r[col("Kernel Name")].split("(")[0], "dram_read": val(r, "dram__bytes_read.sum"),
                     "dram_write": val(r, "dram__bytes_write.sum"), "seconds": val(r, "gpu__time_duration.sum"),
                     "ipc_per_sm": float(r[col("sm__inst_executed.avg.per_cycle_active")]),
                     "threads_per_inst": float(r[col("smsp__thread_inst_executed_per_inst_executed.ratio")]),
                     "alu_pipe_pct": float(r[col("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")]),
                     "fma_pipe_pct": float(r[col("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")]),
                     "l1_hit_pct": float(r[col("l1tex__t_sector_hit_rate.pct")]),
                     "registers": int(float(r[col("launch__registers_per_thread")]))})
# mesh scenes run three kernels per extend round (pass 1 with entries, mesh walk, classification): one "launch" = one round
rounds, cur = [], None
for l in launches:
    first = not ("walk_mesh" in l["kernel"] or "classify_mesh" in l["kernel"])
    if first or cur is None:
        cur = dict(l, kernels=[l["kernel"]], w=l["seconds"])
        for k in ("ipc_per_sm", "threads_per_inst", "alu_pipe_pct", "fma_pipe_pct", "l1_hit_pct"):
            cur[k] = l[k] * l["seconds"]
        rounds.append(cur)
    else:
        cur["kernels"].append(l["kernel"])
        for k in ("dram_read", "dram_write", "seconds"):
            cur[k] += l[k]
        for k in ("ipc_per_sm", "threads_per_inst", "alu_pipe_pct", "fma_pipe_pct", "l1_hit_pct"):
            cur[k] += l[k] * l["seconds"]
        cur["registers"] = max(cur["registers"], l["registers"])
for r_ in rounds:
    for k in ("ipc_per_sm", "threads_per_inst", "alu_pipe_pct", "fma_pipe_pct", "l1_hit_pct"):
        r_[k] /= r_["seconds"]
    r_["kernel"] = " + ".join(dict.fromkeys(r_["kernels"]))
per_kernel = {}
for l in launches:
    a = per_kernel.setdefault(l["kernel"], {"launches": 0, "seconds": 0.0, "lanes_x_s": 0.0, "ipc_x_s": 0.0})
    a["launches"] += 1; a["seconds"] += l["seconds"]; a["lanes_x_s"] += l["threads_per_inst"] * l["seconds"]; a["ipc_x_s"] += l["ipc_per_sm"] * l["seconds"]
launches = rounds
n = len(launches)
tot = sum(l["dram_read"] + l["dram_write"] for l in launches)
secs = sum(l["seconds"] for l in launches)
w = lambda k: sum(l[k] * l["seconds"] for l in launches) / secs
summary = {"source": sys.argv[1].split("/")[-1], "launches": n, "traffic_bytes_per_launch": tot / n,
           "avg_launch_ms_under_ncu": 1e3 * secs / n, "ipc_per_sm_time_weighted": w("ipc_per_sm"),
           "threads_per_inst_time_weighted": w("threads_per_inst"), "alu_pipe_pct_time_weighted": w("alu_pipe_pct"),
           "fma_pipe_pct_time_weighted": w("fma_pipe_pct"), "l1_hit_pct_time_weighted": w("l1_hit_pct"),
           "registers": launches[0]["registers"] if launches else None, "kernel": launches[0]["kernel"] if launches else None,
           "per_kernel": {k: {"launches": v["launches"], "ms_under_ncu": 1e3 * v["seconds"], "lanes_of_32": v["lanes_x_s"] / v["seconds"],
                              "ipc_per_sm": v["ipc_x_s"] / v["seconds"]} for k, v in per_kernel.items()},
           "per_launch": [{"ms": 1e3 * l["seconds"], "dram_MB": (l["dram_read"] + l["dram_write"]) / 1e6} for l in launches]}
json.dump(summary, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k != "per_launch"}, indent=1))
