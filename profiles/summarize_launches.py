"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the step)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        name = re.sub(r"<.*", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / tot * 100:6.1f}%")
    print(f"{'TOTAL':60s} {sum(v[0] for v in agg.values()):8d} {tot / 1e3:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
