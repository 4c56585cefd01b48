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
    if "extend" not in r[col("Kernel Name")]: continue
    launches.append({"kernel": r[col("Kernel Name")].split("(")[0], "dram_read": val(r, "dram__bytes_read.sum"),
                     "dram_write": val(r, "dram__bytes_write.sum"), "seconds": val(r, "gpu__time_duration.sum"),
                     "ipc_per_sm": float(r[col("sm__inst_executed.avg.per_cycle_active")]),
                     "threads_per_inst": float(r[col("smsp__thread_inst_executed_per_inst_executed.ratio")]),
                     "alu_pipe_pct": float(r[col("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")]),
                     "fma_pipe_pct": float(r[col("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")]),
                     "l1_hit_pct": float(r[col("l1tex__t_sector_hit_rate.pct")]),
                     "registers": int(float(r[col("launch__registers_per_thread")]))})
n = len(launches)
tot = sum(l["dram_read"] + l["dram_write"] for l in launches)
secs = sum(l["seconds"] for l in launches)
w = lambda k: sum(l[k] * l["seconds"] for l in launches) / secs
summary = {"source": sys.argv[1].split("/")[-1], "launches": n, "traffic_bytes_per_launch": tot / n,
           "avg_launch_ms_under_ncu": 1e3 * secs / n, "ipc_per_sm_time_weighted": w("ipc_per_sm"),
           "threads_per_inst_time_weighted": w("threads_per_inst"), "alu_pipe_pct_time_weighted": w("alu_pipe_pct"),
           "fma_pipe_pct_time_weighted": w("fma_pipe_pct"), "l1_hit_pct_time_weighted": w("l1_hit_pct"),
           "registers": launches[0]["registers"] if launches else None, "kernel": launches[0]["kernel"] if launches else None,
           "per_launch": [{"ms": 1e3 * l["seconds"], "dram_MB": (l["dram_read"] + l["dram_write"]) / 1e6} for l in launches]}
json.dump(summary, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k != "per_launch"}, indent=1))
