#!/usr/bin/env python
"""Markdown summary of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv`): per kernel (full signature, so the reference's forward and backward renderCUDA<3> stay apart) launches, average
duration, share of the listed kernel time and DRAM traffic / bandwidth per launch.

    python tools/summarize_launches.py gpurun_out/r02_ref_launches.csv [title] > profiles/....md
"""
import collections
import csv
import re
import sys


def short(name):
    m = re.match(r"(?:void )?([\w:<>, ]+?)\(", name)
    base = m.group(1) if m else name[:60]
    base = re.sub(r"radix::policy_hub<[^>]*>", "policy", base)
    if "renderCUDA" in name:   # the reference's two render kernels share a name: tell them apart by signature
        base += " [backward, backward.cu:143]" if "float3 *" in name else " [forward, forward.cu:256]"
    if "preprocessCUDA" in name:
        base += " [backward, backward.cu:581]" if "const float3 *" in name else " [forward, forward.cu:148]"
    return base[:90]


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        key = r["Kernel Name"]
        agg.setdefault(key, collections.defaultdict(list))[r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v["gpu__time_duration.sum"]) for v in agg.values())
    print(f"# {title}\n")
    print("| kernel | launches | avg us | share | DRAM read MB | DRAM write MB | DRAM GB/s |")
    print("|---|---|---|---|---|---|---|")
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        t = v["gpu__time_duration.sum"]
        n = len(t)
        avg_ns = sum(t) / n
        rd, wr = sum(v["dram__bytes_read.sum"]) / n, sum(v["dram__bytes_write.sum"]) / n
        print(f"| `{short(name)}` | {n} | {avg_ns / 1e3:.1f} | {sum(t) / tot * 100:.1f}% | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
              f"{(rd + wr) / avg_ns:.0f} |")
    print(f"\nTotal listed kernel time: {tot / 1e6:.2f} ms over {sum(len(v['gpu__time_duration.sum']) for v in agg.values())} launches.")


if __name__ == "__main__":
    main()
