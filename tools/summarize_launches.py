"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import io
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {v[0]} | {v[1]/1e3:.2f} | {v[1]/v[0]:.1f} | {v[1]/tot*100:.1f}% |")
print(f"| total | {sum(v[0] for v in agg.values())} | {tot/1e3:.2f} | | |")
