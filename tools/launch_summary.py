"""Summarises an ncu launch list (csv with gpu__time_duration.sum): per kernel count, mean, total, share; `--tail N`
restricts to the last N launches (e.g. the steady-state epochs)."""
import csv
import collections
import re
import sys

path = sys.argv[1]
tail = int(sys.argv[sys.argv.index("--tail") + 1]) if "--tail" in sys.argv else 0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u.replace("second", "s").replace("n", "n"), 1.0) if u in ("ns", "us", "ms", "s") else (1e-3 if u.startswith("n") else 1.0)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v))
if tail:
    rows = rows[-tail:]
agg = collections.OrderedDict()
for k, v in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/tot*100:5.1f} %  {n:4d} x {t/n:8.2f} us  {k[:100]}")
