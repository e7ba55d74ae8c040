"""DRAM traffic of one kernel class per forward from an ncu CSV
(--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:<class>).
Usage: python tools/ncu_traffic.py gpurun_out/gemm_traffic.csv <launches per forward> profiles/gemm_traffic_r1.json
Takes the LAST forward in the capture (the earlier ones include warm-up / first-touch effects)."""
import collections
import csv
import json
import sys


def val(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    return v * scale.get(u, 1.0)


def main(path, per_forward, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    by_id = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = by_id.setdefault(row["ID"], {"name": row["Kernel Name"]})
        d[row["Metric Name"]] = val(row)
    launches = list(by_id.values())[-per_forward:]
    rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in launches)
    wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in launches)
    t = sum(l.get("gpu__time_duration.sum", 0.0) for l in launches)
    res = {"launches": len(launches), "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes_per_launch": (rd + wr) / len(launches),
           "duration_s_sum": t, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
                                         "last forward of the capture (bench.py --steps 1 --warmup 3, ViT-B batch 64)"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3])
