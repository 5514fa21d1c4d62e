"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total, count, share."""
import collections
import csv
import sys


def main(path):
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(d["Metric Unit"], v)
        name = d["Kernel Name"].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} us  {v[0]:4d}x  {v[1] / v[0]:10.1f} us/launch  {100 * v[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
