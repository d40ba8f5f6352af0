"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the captured window)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"# {sys.argv[1]}: {sum(n for n, _ in agg.values())} launches, {tot / 1e3:.3f} ms total (ncu-serialised, cold-cache: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / 1e3:10.3f} ms {100 * t / tot:6.2f}%  n={n:4d}  avg={t / n:9.1f} us  {k}")
