"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total device
time and share of the profiled region.  Usage: python tools/ncu_summary.py gpurun_out/launches.csv > profiles/x.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.2f%% |" % (k[:90], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    print("\n%d launches, %.2f ms device time in total (cold-cache, serialised by ncu: compare shares, not absolutes)" % (n, tot / 1e3))


if __name__ == "__main__":
    main(sys.argv[1])
