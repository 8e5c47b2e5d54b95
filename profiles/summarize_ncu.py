"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram bytes]) per kernel and per pass.
usage: python profiles/summarize_ncu.py file.csv [launches_per_pass]"""
import csv, sys, collections

path = sys.argv[1]
per_pass = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
by_id = collections.OrderedDict()
for r in rd:
    i = r["ID"]
    e = by_id.setdefault(i, {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        e["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    elif m == "dram__bytes_read.sum":
        e["rd"] = v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    elif m == "dram__bytes_write.sum":
        e["wr"] = v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
launches = list(by_id.values())
if per_pass:
    last = launches[-per_pass:]
    print(f"last pass ({per_pass} launches):")
    tot = 0.0
    for e in last:
        tot += e["us"]
        extra = f" rd={e['rd']/1e6:8.2f}MB wr={e['wr']/1e6:7.2f}MB" if "rd" in e else ""
        print(f"  {e['name'][:70]:70s} {e['us']:9.2f} us{extra}")
    print(f"  total {tot:.1f} us")
    for e in last:
        print(f"  share {100*e['us']/tot:5.1f}%  {e['name'][:60]}")
agg = collections.OrderedDict()
for e in launches:
    a = agg.setdefault(e["name"], [])
    a.append(e["us"])
print("per kernel over the whole capture:")
for k, v in agg.items():
    print(f"  {k[:70]:70s} n={len(v):4d} mean={sum(v)/len(v):9.2f} us min={min(v):9.2f}")
