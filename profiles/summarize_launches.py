"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python profiles/summarize_launches.py gpurun_out/launches_global.csv > profiles/rNN_launches_<workload>.md
"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"source: {path}  ({len(rows) - 1} launches, {tot / 1e3:.3f} ms of kernel time; per-launch times are "
          "cold-cache and serialised under ncu: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / tot:.4f} |")


if __name__ == "__main__":
    main(sys.argv[1])
