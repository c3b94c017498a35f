import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"total {tot:.3f} ms over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
    print(f"{t:9.3f} ms {100*t/tot:5.1f}%  n={n:4d}  avg {1e3*t/n:8.1f} us  {k}")
