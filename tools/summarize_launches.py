#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: summarize_launches.py launches.csv [steps_in_capture]   (the last 1/steps of the launches = one step)"""
import csv
import re
import sys
from collections import defaultdict


def main(path, steps=1):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
    rows.sort()
    n = len(rows)
    per = n // steps
    last = rows[n - per:]
    agg = defaultdict(lambda: [0, 0.0])
    for _, name, ns in last:
        short = re.sub(r"<.*", "", name)
        short = re.sub(r"\(.*", "", short)
        agg[short][0] += 1
        agg[short][1] += ns
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "mrb" in k or k.startswith(("conv_tc", "conv_wgrad", "conv_prepare", "roi_align", "nms_", "focal", "dcn_", "psroi", "roi_pool")))
    print("launches in capture: %d; analysed (last step): %d; device time of the step: %.2f ms (serialised, cold)" % (n, per, tot / 1e6))
    print("share of libmrb_b200.so kernels: %.1f%%" % (100 * ours / tot))
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("| %s | %d | %.1f | %.1f%% |" % (k[:90], c, ns / 1e3, 100 * ns / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
