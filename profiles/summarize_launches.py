"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [n_steps]"""
import collections
import csv
import re
import sys


def main(path, steps=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        v = float(x["Metric Value"].replace(",", ""))
        unit = x["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'total us':>12} {'count':>6} {'avg us':>10} {'share':>7}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} {v[0]:6d} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%  {k[:80]}")
    print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches" +
          (f" = {tot / steps:.1f} us per step" if steps else ""))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
