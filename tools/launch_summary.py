"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (profiles/*.md)."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), float(r["Metric Value"]) * (1e-3 if r["Metric Unit"] == "ns" else 1.0)))
agg = collections.OrderedDict()
for k, us in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, n, us, us / n, 100 * us / tot))
print("\ntotal %.3f ms over %d launches" % (tot / 1e3, len(rows)))
