"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv` output.
Usage: ncu -i x.ncu-rep --page source --csv --kernel-id ::regex:NAME:N > src.csv; python tools/ncu_source_top.py src.csv [K]"""
import csv
import sys


def main(path, k=30):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("total samples", tot, "instructions", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
    print("stall totals:", sorted(((v, s) for s, v in agg.items() if v), reverse=True)[:8])
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:k]:
        st = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
        print(r[ix["# Samples"]].rjust(6), r[ix["Source"]].strip()[:64].ljust(64), st,
              "L2sec %s/%s" % (r[ix["L2 Theoretical Sectors Global"]], r[ix["L2 Theoretical Sectors Global Ideal"]]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
