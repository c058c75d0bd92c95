"""Summarise an ncu launch-list CSV (gpu__time_duration.sum) by kernel."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] == "ns" else v
    k = r[ki].replace("ub::", "").replace("<unnamed>::", "").split("(")[0][:48]
    d[k][0] += 1; d[k][1] += v
tot = sum(v[1] for v in d.values())
print("total %.1f us over %d launches" % (tot, sum(v[0] for v in d.values())))
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("%-50s n=%3d total %9.1f us %5.1f%% avg %8.1f us" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
