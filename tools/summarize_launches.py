"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name + grid."""
import csv, collections, sys
path = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
rows = rows[int(len(rows) * frac):]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    ns = v * 1e3 if u == "us" else v * 1e6 if u == "ms" else v
    k = r["Kernel Name"].split("(")[0][-42:] + " grid=" + r["Grid Size"]
    agg[k][0] += 1; agg[k][1] += ns
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%-72s n=%4d total=%8.3f ms share=%5.1f%% avg=%8.1f us" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3))
print("total %.3f ms over %d launches" % (tot / 1e6, len(rows)))
