"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean time, share."""
import collections
import csv
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    data = data[skip:]
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v /= 1000.0
        elif r[ui] in ("ms", "msecond"):
            v *= 1000.0
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"| kernel | launches | mean us | share |\n|---|---:|---:|---:|")
    for k, v in agg.items():
        print(f"| `{k[:70]}` | {len(v)} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.1f}% |")
    print(f"| total | {sum(len(v) for v in agg.values())} | {tot:.1f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
