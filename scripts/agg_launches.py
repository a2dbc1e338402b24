"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void |<unnamed>::|at::native::|at::", "", name)[:100]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total %.1f us in %d launches" % (tot, sum(n for n, _ in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%10.1f us %5d x %5.1f%%  %s" % (t, n, 100 * t / tot, k))
