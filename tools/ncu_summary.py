#!/usr/bin/env python
"""Reduce `ncu -i X.ncu-rep --page raw --csv` output to the handful of metrics quoted in profiles/: per kernel the duration,
DRAM bytes read + written, L2 / SM throughput (% of peak), achieved occupancy, registers, tensor-pipe activity.
usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py > summary.json"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__pipe_tensor_subpipe_tc_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
}


def main():
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    if len(rows) < 3:
        print("{}")
        return
    head, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(head)}
    out = []
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        rec = {"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]])[:80]}
        for m, name in WANT.items():
            if m in idx:
                try:
                    v = float(r[idx[m]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[m]]
                if name == "duration_us":
                    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
                if name.endswith("_MB"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                rec[name] = round(v, 3)
        out.append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
