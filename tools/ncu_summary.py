#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel) into the handful of numbers DESIGN.md / profiles/ quote.
usage: ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 sectors global ld"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slot util %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        ix = {h: i for i, h in enumerate(hdr)}
        print(f"== {rep}: {vals[ix['Kernel Name']] if 'Kernel Name' in ix else ''}")
        for k, name in KEYS:
            if k in ix:
                print(f"  {name:28s} {vals[ix[k]]:>18s} {units[ix[k]]}")
        st = [(float(vals[i]), h.split('issue_stalled_')[1].split('_per_')[0]) for h, i in ix.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))


if __name__ == "__main__":
    main()
