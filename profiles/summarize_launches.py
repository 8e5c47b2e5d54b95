"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: summarize_launches.py file.csv [last_n]"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("nsecond", "ns") else (v * 1e3 if unit in ("msecond", "ms") else v)
        seq.append((row["Kernel Name"][:70], v))
tail = seq[-int(sys.argv[2]):] if len(sys.argv) > 2 else seq
agg = collections.OrderedDict()
for k, v in tail:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(tail)} launches, {tot / 1e3:.2f} ms in kernels (serialised, cold-cache per-launch times)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:72s} n={n:5d} mean={t / n:8.2f} us share={100 * t / tot:5.1f}%")
