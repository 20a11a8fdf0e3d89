#!/usr/bin/env python
"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` (one column per kernel instance)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput", "lts__t_bytes.sum",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "issue_stalled", "l1tex__data_bank_conflicts",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block", "sm__ctas_launched",
        "achieved_occupancy", "sm__maximum_warps", "smsp__warps_eligible", "lts__t_sector_hit_rate"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for inst in rows[2:]:
        print("==", inst[hdr.index("Kernel Name")][:110], inst[hdr.index("Grid Size")], inst[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS + extra):
                print(f"   {h} [{units[i]}] = {inst[i]}")


if __name__ == "__main__":
    main()
