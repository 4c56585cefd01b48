"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python tools/launch_summary.py file.csv [--last-step]
--last-step keeps only the launches between the last two resolve_kernel launches (= the last render step)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr = r; start = i + 1; break
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
body = [r for r in rows[start:] if len(r) > vi]
if "--last-step" in sys.argv:
    res = [i for i, r in enumerate(body) if r[ki].startswith("resolve_kernel") or "resolve_kernel" in r[ki]]
    if len(res) >= 2: body = body[res[-2] + 1: res[-1] + 1]
    print(f"last render step: {len(body)} launches")
agg = collections.OrderedDict()
for r in body:
    if len(r) <= vi: continue
    name = r[ki].split('(')[0]; v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} n={n:4d} {t:10.1f} us {100*t/tot:5.1f}%")
print(f"total {tot:.1f} us (cold-cache, serialised: compare shares)")
