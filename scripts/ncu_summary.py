#!/usr/bin/env python
"""Condense an ncu --set full report into the few counters DESIGN.md / bench.py quote.

    python scripts/ncu_summary.py gpurun_out/prof_<kernel>.ncu-rep > profiles/rNN_ncu_<kernel>.txt
"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct"
    r"|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)"
    r"|sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active"
    r"|sm__inst_executed_pipe_(xu|fma|alu|lsu|tmem|uniform)\.avg\.pct_of_peak_sustained_active"
    r"|sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active"
    r"|sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active"
    r"|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active"
    r"|sm__cycles_active\.avg|sm__cycles_elapsed\.max"
    r"|launch__(registers_per_thread|grid_size|block_size|cluster_size|shared_mem_per_block_dynamic|waves_per_multiprocessor|occupancy_limit_\w+)"
    r"|smsp__sass_inst_executed_op_tmem_(ldt|stt)\.sum|l1tex__data_pipe_tc_wavefronts_mem_shared\.sum"
    r"|l1tex__data_bank_conflicts_pipe_lsu\.sum|smsp__inst_executed\.sum"
    r"|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio)$")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel: %s" % r[ki][:160])
        stalls = []
        for h, u, v in zip(hdr, units, r):
            if not KEEP.match(h):
                continue
            if h.startswith("smsp__average_warps_issue_stalled"):
                try:
                    stalls.append((float(v.replace(",", "")), h.split("stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
                continue
            print("  %-82s %-16s %s" % (h, u, v))
        if stalls:
            print("  top stall reasons (warps stalled per issue-active cycle):")
            for v, n in sorted(stalls, reverse=True)[:6]:
                print("    %-40s %.3f" % (n, v))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
