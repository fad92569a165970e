"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit == "ns" else v * 1000.0 if unit == "ms" else v
        k = row["Kernel Name"].split("(")[0][:48]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':50s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:50s} {v[0]:5d} {v[1]:10.1f} {v[1] / v[0]:9.2f} {100 * v[1] / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
