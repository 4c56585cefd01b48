"""Per-source-line aggregation of an `ncu --import-source on` report:
    python tools/ncu_lines.py rep.ncu-rep <launch index> [top N]
Runs `ncu -i rep --page source --csv --print-source cuda,sass` and sums, per CUDA source line, the warp-level
instructions executed, thread-level instructions executed and stall samples of the SASS under it."""
import csv, subprocess, sys, io, collections
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr = None, None
agg = collections.OrderedDict()
cur = None
kern = ""
for r in rows:
    if not r: continue
    if r[0] == "Function Name": kern = r[1]; continue
    if r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip())
        agg.setdefault(cur, [0, 0, 0])
        continue
    if cur is None or r[2] == "...": continue
    try:
        ie, te, sm = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("Thread Instructions Executed")]), int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    a = agg[cur]; a[0] += ie; a[1] += te; a[2] += sm
tot = [sum(a[i] for a in agg.values()) for i in range(3)]
print(kern[:100])
print(f"total warp-inst {tot[0]}  thread-inst {tot[1]}  avg lanes {tot[1]/max(tot[0],1):.1f}  samples {tot[2]}")
items = sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]
for (f, ln, src), (ie, te, sm) in items:
    print(f"{100*sm/max(tot[2],1):5.1f}% smp {100*ie/max(tot[0],1):5.1f}% inst lanes {te/max(ie,1):5.1f}  {f}:{ln}: {src[:110]}")
