"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit.startswith("us"): v *= 1e3
    elif unit.startswith("ms"): v *= 1e6
    elif unit.startswith("s") and not unit.startswith("ns"): v *= 1e9
    nm = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((nm, v, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(v for _, v, _, _ in rows)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for nm, v, g, b in rows:
    a = agg[nm]; a[0] += 1; a[1] += v; a[2] = max(a[2], v)
print(f"total {tot/1e6:.3f} ms over {len(rows)} launches")
print(f"{'kernel':60s} {'n':>6s} {'total ms':>10s} {'share':>7s} {'avg us':>9s} {'max us':>9s}")
for nm, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{nm[:60]:60s} {n:6d} {t/1e6:10.3f} {100*t/tot:6.1f}% {t/n/1e3:9.1f} {mx/1e3:9.1f}")
if len(sys.argv) > 2:
    k = sys.argv[2]
    print("\nlaunches of", k)
    for nm, v, g, b in rows:
        if k in nm: print(f"   {v/1e3:10.1f} us grid={g} block={b}")
