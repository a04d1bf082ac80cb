#!/usr/bin/env python
"""Read an `ncu --set full` report (via `ncu -i X.ncu-rep --page raw --csv`) and print the handful of counters the
roofline / DESIGN notes cite, one block per profiled launch.  Usage: python profiles/ncu_extract.py X.ncu-rep [more keys]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__inst_executed_pipe_uniform",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    keys = KEYS + sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {d['Kernel Name'][:110]}  grid {d['Grid Size']} block {d['Block Size']}")
        for i, h in enumerate(hdr):
            if any(k in h for k in keys):
                print(f"{h:90s} {r[i]:>18s} {units[i]}")
        print()


if __name__ == "__main__":
    main()
