"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import collections
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("%-60s %8s %12s %8s %10s" % ("kernel", "launches", "total_us", "share", "avg_us"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-60s %8d %12.1f %7.1f%% %10.2f" % (k[:60], cnt[k], v, 100 * v / T, v / cnt[k]))
print("%-60s %8d %12.1f" % ("TOTAL", sum(cnt.values()), T))
