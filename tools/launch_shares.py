"""Per-kernel time shares, busy lanes and IPC from an ncu --csv launch list (gpu__time_duration.sum,
smsp__thread_inst_executed_per_inst_executed.ratio, sm__inst_executed.avg.per_cycle_active)."""
import csv, collections, sys
for fn in sys.argv[1:]:
    rows = list(csv.reader(open(fn)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        per.setdefault((r[0], r[kn].split("(")[0]), {})[r[mn]] = (float(r[mv].replace(",", "")), r[mu])
    agg = collections.OrderedDict()
    for (i, k), m in per.items():
        d, u = m["gpu__time_duration.sum"]
        d = d / 1e6 if u in ("ns", "nsecond") else d / 1e3 if u in ("us", "usecond") else d
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += d
        a[2] += m["smsp__thread_inst_executed_per_inst_executed.ratio"][0] * d; a[3] += m["sm__inst_executed.avg.per_cycle_active"][0] * d
    tot = sum(a[1] for a in agg.values())
    print(fn, "total ms", round(tot, 2))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k[:60]:60s} n={a[0]:3d} {a[1]:8.2f} ms {100 * a[1] / tot:5.1f}%  lanes {a[2] / a[1]:5.1f} ipc {a[3] / a[1]:4.2f}")
