#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total / average
duration and share of the captured window.  Usage: python profiles/summarize_launches.py launches.csv [launches_per_step]"""
import collections
import csv
import re
import sys


def main():
    lines = open(sys.argv[1]).read().splitlines()
    start = [n for n, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
    agg = collections.OrderedDict()
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"launches captured: {len(rows)}   total device time: {tot/1e3:.1f} us" + (f"   ({len(rows)/per_step:.2f} steps of {per_step} launches)" if per_step else ""))
    print(f"{'total us':>10s} {'share':>6s} {'n':>5s} {'avg us':>8s}  kernel")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1]/1e3:10.1f} {v[1]/tot*100:5.1f}% {v[0]:5d} {v[1]/v[0]/1e3:8.1f}  {k}")
    cls = collections.OrderedDict([("tcgen05 GEMM (k_gemm_tc)", 0.0), ("flash attention (fwd+bwd+delta)", 0.0), ("bandwidth-bound (norm/act/CE/optimizer/gather/rope/colsum/...)", 0.0)])
    for k, v in agg.items():
        if "k_gemm_tc" in k:
            cls["tcgen05 GEMM (k_gemm_tc)"] += v[1]
        elif "flash" in k or "attn" in k:
            cls["flash attention (fwd+bwd+delta)"] += v[1]
        else:
            cls["bandwidth-bound (norm/act/CE/optimizer/gather/rope/colsum/...)"] += v[1]
    print()
    for k, v in cls.items():
        print(f"{v/1e3:10.1f} us {v/tot*100:5.1f}%  {k}")


if __name__ == "__main__":
    main()
