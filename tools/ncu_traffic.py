"""DRAM traffic of one kernel class per forward from an ncu CSV
(--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:<class>).
Usage: python tools/ncu_traffic.py gpurun_out/gemm_traffic.csv <launches per forward> profiles/gemm_traffic_r1.json [forward index]
Takes the LAST forward in the capture (the earlier ones include warm-up / first-touch effects), or -- when the capture continues
past the batch-64 steps (the bench's batch-2 latency leg) -- the forward with the given 0-based index."""
import collections
import csv
import json
import sys


def val(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    return v * scale.get(u, 1.0)


def main(path, per_forward, out, index=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    by_id = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = by_id.setdefault(row["ID"], {"name": row["Kernel Name"]})
        d[row["Metric Name"]] = val(row)
    allv = list(by_id.values())
    launches = allv[-per_forward:] if index is None else allv[index * per_forward:(index + 1) * per_forward]
    rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in launches)
    wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in launches)
    t = sum(l.get("gpu__time_duration.sum", 0.0) for l in launches)
    res = {"launches": len(launches), "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes_per_launch": (rd + wr) / len(launches),
           "duration_s_sum": t, "kernels": sorted(set(l["name"].split("(")[0][-48:] for l in launches)), "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
                                         "last forward of the capture (bench.py --steps 1 --warmup 3, ViT-B batch 64)"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else None)
