"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/summarize_launches.py <csv>"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if v != v:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)      # -> us
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'share':>7} {'ms':>9} {'launches':>8}  kernel")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1] / tot * 100:6.2f}% {v[1] / 1e3:9.2f} {v[0]:8d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
