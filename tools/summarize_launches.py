"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel count, mean time, share of the serialised kernel time and (when captured) DRAM bytes per launch."""
import collections
import csv
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, mi, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    agg = collections.OrderedDict()
    for r in data:
        if int(r[ii]) < skip:
            continue
        v = float(r[vi].replace(",", ""))
        name = r[ki].split("(")[0].replace("mwe::", "")
        if r[mi].startswith("gpu__time"):
            v = v / 1000.0 if r[ui] in ("ns", "nsecond") else (v * 1000.0 if r[ui] in ("ms", "msecond") else v)
            agg.setdefault(name, {"t": [], "b": 0.0})["t"].append(v)
        elif r[mi].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
            agg.setdefault(name, {"t": [], "b": 0.0})["b"] += v
    tot = sum(sum(v["t"]) for v in agg.values())
    has_b = any(v["b"] for v in agg.values())
    print("| kernel | launches | mean us | share |" + (" DRAM MB / launch |" if has_b else ""))
    print("|---|---:|---:|---:|" + ("---:|" if has_b else ""))
    for k, v in agg.items():
        n = len(v["t"])
        line = f"| `{k[:70]}` | {n} | {sum(v['t']) / n:.2f} | {100 * sum(v['t']) / tot:.1f}% |"
        if has_b:
            line += f" {v['b'] / n / 1e6:.1f} |"
        print(line)
    print(f"| total | {sum(len(v['t']) for v in agg.values())} | {tot:.1f} | 100% |" + (f" {sum(v['b'] for v in agg.values()) / 1e6:.0f} (sum) |" if has_b else ""))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
