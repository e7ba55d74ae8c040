"""Top source lines by stall samples from `ncu -i rep --page source --csv --print-source cuda,sass > x.csv`.
Usage: python tools/ncu_lines_top.py x.csv [K]"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(path, k=30):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hi]
    si = hdr.index("# Samples")
    ei = hdr.index("Instructions Executed")
    stallcols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    lines = [r for r in rows[hi + 1:] if r and r[0].isdigit()]
    print("total samples", sum(num(r[si]) for r in lines))
    for r in sorted(lines, key=lambda r: -num(r[si]))[:k]:
        st = sorted(((num(r[i]), hdr[i][6:]) for i in stallcols), reverse=True)[:3]
        print(r[0].rjust(4), r[si].rjust(6), r[ei].rjust(9), r[1].strip()[:90].ljust(90), st)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
