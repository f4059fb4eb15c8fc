#!/usr/bin/env python3
"""Launch-share table from an `ncu --metrics gpu__time_duration.sum --csv` log:
python tools/launch_shares.py launches.csv "title" > profiles/rNN_launch_shares.md"""
import csv
import glob
import os
import re
import sys


def own_kernels():
    """names of the __global__ functions of normalisr_b200/csrc (everything else in a log is torch's)"""
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "normalisr_b200", "csrc")
    names = set()
    for f in glob.glob(os.path.join(root, "*.cu")):
        src = open(f).read()
        for m in re.finditer(r"__global__", src):
            head = src[m.end():m.end() + 400]
            head = re.sub(r"__(launch_bounds|cluster_dims)__\s*\([^)]*\)", " ", head)
            k = re.search(r"\b(\w+)\s*\(", head)
            if k:
                names.add(k.group(1))
    return names


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
    rows = [r for r in csv.reader(line for line in open(path) if line.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k_name, k_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    order, stats = [], {}
    own = own_kernels()
    for r in rows:
        name = re.sub(r"^(void )?(<unnamed>::)?", "", r[k_name])
        name = re.sub(r"\(.*$", "", name)
        if name.split("<")[0] not in own:
            continue                         # torch kernels (synthetic data generation, glue)
        ns = float(r[k_val].replace(",", ""))
        if name not in stats:
            stats[name] = [0, 0.0]
            order.append(name)
        stats[name][0] += 1
        stats[name][1] += ns
    total = sum(v[1] for v in stats.values())
    print("# %s\n" % title)
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (serialised, cold cache: compare shares, not times).\n")
    print("| kernel | launches | mean ms | share of device time |\n|---|---|---|---|")
    for name in order:
        c, ns = stats[name]
        print("| `%s` | %d | %.3f | %.1f %% |" % (name, c, ns / c / 1e6, 100 * ns / total))
    print("\ntotal device time of these launches: %.1f ms" % (total / 1e6))


if __name__ == "__main__":
    main()
