"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel: python scripts/launch_summary.py x.csv [first last]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    if len(sys.argv) > 3:
        rows = rows[int(sys.argv[2]):int(sys.argv[3])]
    agg = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("thmr::", "").replace("void ", "")
        name = re.sub(r"\(int\)", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit.startswith("ms") else 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:90]}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
    print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | 100% |")


if __name__ == "__main__":
    main()
